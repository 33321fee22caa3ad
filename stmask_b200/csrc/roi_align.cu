// RoIAlign forward (average pooling) on NHWC feature maps — the step right after the temporal-fusion concat:
// bbox_feat_extractor crops the 633-channel concat (or its padded 640-channel B200 layout) for every candidate
// box (reference layers/modules/track_to_segment_head.py:65-88 -> mmcv.ops.roi_align, output 7x7,
// spatial_scale 1, adaptive sampling grid, aligned = True).
//
//   out[r, c, i, j] = mean over the gh x gw sample points of bin (i, j) of bilinear(feat[b_r, :, :, c], y, x)
//
// One CTA per (roi, output row); threads across (bin, 8-channel chunk) (channels-last => every corner of every sample point is
// one contiguous C-vector: 16-byte loads when C % 8 == 0 in bf16, else element-wise), fp32 accumulation.
// HBM/L2-bound by construction: every corner vector is read once per sample point; the maps are tiny (L2-resident).
#include <type_traits>

#include "common.cuh"

namespace stm {
namespace {

struct RoiArgs {
  StmRoiAlignDesc d;
  const void* feat;
  const float* rois;
  void* out;
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// RoIAlign edge rule (Detectron2 / torchvision roi_align bilinear_interpolate): a point outside [-1, H] x [-1, W]
// contributes nothing; otherwise clamp to [0, size - 1]
struct Corner4 {
  float w[4];
  int64_t o[4];
  bool live;
};
__device__ __forceinline__ Corner4 corners(float y, float x, int H, int W, int64_t sh, int64_t sw) {
  Corner4 c;
  c.live = !(y < -1.f || y > (float)H || x < -1.f || x > (float)W);
  y = fmaxf(y, 0.f);
  x = fmaxf(x, 0.f);
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  c.w[0] = hy * hx; c.w[1] = hy * lx; c.w[2] = ly * hx; c.w[3] = ly * lx;
  c.o[0] = yl * sh + xl * sw; c.o[1] = yl * sh + xh * sw; c.o[2] = yh * sh + xl * sw; c.o[3] = yh * sh + xh * sw;
  return c;
}

template <typename T, typename OT, bool VEC8>
__global__ void __launch_bounds__(256) roi_align_kernel(const RoiArgs a) {
  const StmRoiAlignDesc& d = a.d;
  const int pi = blockIdx.x % d.pooled_h;          // one CTA per (roi, output row): pooled_w bins x channel chunks
  const int r = blockIdx.x / d.pooled_h;
  const float* roi = a.rois + (size_t)r * 5;
  const int b = (int)roi[0];
  const float off = d.aligned ? 0.5f : 0.f;
  const float x1 = roi[1] * d.spatial_scale - off, y1 = roi[2] * d.spatial_scale - off;
  const float x2 = roi[3] * d.spatial_scale - off, y2 = roi[4] * d.spatial_scale - off;
  float rw = x2 - x1, rh = y2 - y1;
  if (!d.aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bh = rh / d.pooled_h, bw = rw / d.pooled_w;
  const int gh = d.sampling_ratio > 0 ? d.sampling_ratio : (int)ceilf(rh / d.pooled_h);
  const int gw = d.sampling_ratio > 0 ? d.sampling_ratio : (int)ceilf(rw / d.pooled_w);
  const float inv = 1.f / (float)max(gh * gw, 1);
  const bool b_ok = b >= 0 && b < d.batch;
  const T* img = reinterpret_cast<const T*>(a.feat) + (int64_t)(b_ok ? b : 0) * d.feat_stride_n;
  constexpr int CPT = VEC8 ? 8 : 1;
  const int chunks = (d.c + CPT - 1) / CPT;
  const bool vec_store = VEC8 && sizeof(OT) == 2 && d.out_stride_c == 1 &&
                         ((d.out_stride_n | d.out_stride_h | d.out_stride_w) & 7) == 0 && (((uintptr_t)a.out) & 15) == 0;
  for (int item = threadIdx.x; item < d.pooled_w * chunks; item += blockDim.x) {
    const int pj = item / chunks, c0 = (item - pj * chunks) * CPT;
    float acc[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
    for (int iy = 0; iy < gh && b_ok; ++iy) {
      const float yy = y1 + pi * bh + (iy + 0.5f) * bh / gh;
      for (int ix = 0; ix < gw; ++ix) {
        const float xx = x1 + pj * bw + (ix + 0.5f) * bw / gw;
        const Corner4 k = corners(yy, xx, d.h, d.w, d.feat_stride_h, d.feat_stride_w);
        if (!k.live) continue;
        if (VEC8) {
          uint4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const uint4*>(img + k.o[q] + c0));
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
            unpack8(v[q], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(k.w[q], f[i], acc[i]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[0] = fmaf(k.w[q], to_f32(img[k.o[q] + c0]), acc[0]);
        }
      }
    }
    OT* orow = reinterpret_cast<OT*>(a.out) + r * d.out_stride_n + pi * d.out_stride_h + pj * d.out_stride_w;
    if (VEC8 && vec_store) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
        w[i] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(orow + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
      for (int i = 0; i < CPT; ++i) orow[(int64_t)(c0 + i) * d.out_stride_c] = from_f32<OT>(acc[i] * inv);
    }
  }
}

template <typename T, typename OT>
int launch_typed(const RoiArgs& args, cudaStream_t s) {
  const StmRoiAlignDesc& d = args.d;
  const unsigned grid = (unsigned)(d.n_rois * d.pooled_h);
  const bool vec8 = std::is_same<T, __nv_bfloat16>::value && (d.c & 7) == 0 && (((uintptr_t)args.feat) & 15) == 0 &&
                    ((d.feat_stride_n | d.feat_stride_h | d.feat_stride_w) & 7) == 0;
  if (vec8) roi_align_kernel<T, OT, std::is_same<T, __nv_bfloat16>::value><<<grid, 256, 0, s>>>(args);
  else roi_align_kernel<T, OT, false><<<grid, 256, 0, s>>>(args);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace

int launch_roi_align(const StmRoiAlignDesc& d, const void* feat, const float* rois, void* out, cudaStream_t stream) {
  RoiArgs args;
  args.d = d; args.feat = feat; args.rois = rois; args.out = out;
  if (d.dtype == STM_BF16) {
    if (d.out_dtype == STM_BF16) return launch_typed<__nv_bfloat16, __nv_bfloat16>(args, stream);
    return launch_typed<__nv_bfloat16, float>(args, stream);
  }
  if (d.out_dtype == STM_BF16) return launch_typed<float, __nv_bfloat16>(args, stream);
  return launch_typed<float, float>(args, stream);
}

}  // namespace stm
