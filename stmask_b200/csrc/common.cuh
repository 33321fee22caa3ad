// Shared device/host helpers for libstmask_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stmask_b200.h"

namespace stm {

// ----------------------------------------------------------------------------------------
// error reporting (thread-local message behind stm_last_error)
// ----------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void clear_error();
void count_launch(int n = 1);

// SM count of the CURRENT device (cached per device ordinal; thread-safe).
int device_sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember, per device
// ordinal, the largest value already set for one kernel instantiation (one cache object per instantiation).
struct SmemAttrCache {
  int v[64];
};
int current_device_ordinal();

#define STM_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::stm::set_error(__VA_ARGS__);             \
      return STM_ERR_INVALID_ARGUMENT;           \
    }                                            \
  } while (0)

#define STM_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::stm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return STM_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// ----------------------------------------------------------------------------------------
// device-side problem descriptions (kernel parameters, passed by value)
// ----------------------------------------------------------------------------------------
struct DcnProblemDev {
  const void* x;
  const void* offset;
  const void* mask;
  void* y;
  int32_t batch, in_h, in_w, out_h, out_w;
  int32_t m_total;      // batch * out_h * out_w output pixels (GEMM rows)
  int32_t tile_begin;   // index of this problem's first M tile in the launch
  int32_t patch;        // tcgen05 kernel: GEMM rows enumerate the map in 8x8 pixel patches instead of raster order
  int64_t x_sn, x_sh, x_sw;
  int64_t y_sn, y_sh, y_sw;
  int64_t off_sn, off_sc, off_sh, off_sw;
  int64_t mask_sn, mask_sc, mask_sh, mask_sw;
};

struct DcnParams {
  DcnProblemDev prob[STM_DCN_MAX_PROBLEMS];
  int32_t n_probs;
  int32_t total_m_tiles;
  int32_t in_c, out_c, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg;
  int32_t flags;
  const void* w;       // OHWI packed
  const float* bias;   // may be null
  const float* fcb_w;  // FCB(ada): conv_offset weight [dg * 2 * kh * kw][4]; null otherwise
};

// ----------------------------------------------------------------------------------------
// scalar conversions
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }

// Bilinear corner weights / element offsets of one deformable sample (DCN border rule,
// SURVEY.md §8b): sample is 0 outside (-1, H) x (-1, W); corners outside the map contribute 0.
// Offsets are relative to the (b, 0, 0, 0) element of an NHWC tensor; invalid corners get
// weight 0 and offset 0.
struct Sample4 {
  float w[4];
  int32_t o[4];
};

__device__ __forceinline__ Sample4 make_sample(float h, float w, int H, int W, int64_t sh, int64_t sw, float scale) {
  Sample4 s;
#pragma unroll
  for (int i = 0; i < 4; ++i) { s.w[i] = 0.f; s.o[i] = 0; }
  if (h > -1.f && w > -1.f && h < (float)H && w < (float)W) {
    const float hf = floorf(h), wf = floorf(w);
    const int h0 = (int)hf, w0 = (int)wf;
    const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
    const bool y0 = h0 >= 0, y1 = h0 + 1 <= H - 1, x0 = w0 >= 0, x1 = w0 + 1 <= W - 1;
    if (y0 && x0) { s.w[0] = hh * hw * scale; s.o[0] = (int32_t)(h0 * sh + w0 * sw); }
    if (y0 && x1) { s.w[1] = hh * lw * scale; s.o[1] = (int32_t)(h0 * sh + (w0 + 1) * sw); }
    if (y1 && x0) { s.w[2] = lh * hw * scale; s.o[2] = (int32_t)((h0 + 1) * sh + w0 * sw); }
    if (y1 && x1) { s.w[3] = lh * lw * scale; s.o[3] = (int32_t)((h0 + 1) * sh + (w0 + 1) * sw); }
  }
  return s;
}

template <typename K>
int ensure_dynamic_smem(K kernel, int bytes, SmemAttrCache& cache) {
  const int dev = current_device_ordinal();
  if (dev < 0 || dev >= 64 || __atomic_load_n(&cache.v[dev], __ATOMIC_ACQUIRE) < bytes) {
    STM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) __atomic_store_n(&cache.v[dev], bytes, __ATOMIC_RELEASE);
  }
  return STM_OK;
}

// host-side launchers implemented in the .cu files
int launch_dcn_simt(const DcnParams& p, int dtype, int offset_dtype, cudaStream_t stream);
int launch_dcn_tc(const StmDcnConv* conv, const DcnParams& p, void* workspace, size_t ws_bytes, cudaStream_t stream);
bool dcn_tc_supported(const StmDcnConv* conv, const StmDcnProblem* probs, int n, const char** why);
bool dcn_tc_shape_supported(const StmDcnConv* conv, const StmDcnProblem* probs, int n, const char** why);
size_t dcn_tc_workspace(const StmDcnConv* conv, const StmDcnProblem* probs, int n);
int dcn_tc_variant(const StmDcnConv* conv, const DcnParams& p, char* buf, size_t len);
// plain convolution with TMA-loaded, shifted-view A operands (conv_tma.cu)
bool conv_tma_shape_supported(const StmDcnConv* conv, const DcnParams& p, const char** why);
int conv_tma_variant(const StmDcnConv* conv, const DcnParams& p, char* buf, size_t len);
int launch_conv_tma(const StmDcnConv* conv, const DcnParams& p, cudaStream_t stream);

int launch_corr_simt(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                     cudaStream_t stream);
int launch_corr_tc(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                   cudaStream_t stream);
int launch_corr_tc_multi(const StmCorrDesc* descs, const void* const* x1s, const void* const* x2s, const void* fa, const void* fb,
                         void* const* outs, int n, cudaStream_t stream);
bool corr_tc_supported(const StmCorrDesc& d, const char** why);
int launch_pool_fc(const void* x, int dtype, int n, int hw, int c, int64_t x_stride_n, int64_t x_stride_p, const float* w,
                   const float* b, int out_features, float* y, cudaStream_t stream);
int launch_detect_nms(const float* conf, const float* loc, const float* centerness, const float* priors, int frames, int P, int C,
                      int top_k, float conf_thresh, float nms_thresh, int32_t* count, int32_t* index, int32_t* cls, float* score,
                      float* box, cudaStream_t stream);
int launch_mask_assembly(const float* proto, const float* coeff, const float* boxes, const int32_t* count, float* masks,
                         uint32_t* bits, int frames, int h, int w, int k, int max_n, cudaStream_t stream);
int launch_mask_iou(const uint32_t* a, const uint32_t* b, const int32_t* na, const int32_t* nb, float* iou, int frames, int max_a,
                    int max_b, int words, cudaStream_t stream);
int launch_track_update(const StmTrackParams& p, const StmTrackState& st, const StmTrackDets& det, const float* mask_iou,
                        const uint8_t* is_first, int32_t* det_slot, uint8_t* keep, cudaStream_t stream);
int launch_roi_align(const StmRoiAlignDesc& d, const void* feat, const float* rois, void* out, cudaStream_t stream);

}  // namespace stm
