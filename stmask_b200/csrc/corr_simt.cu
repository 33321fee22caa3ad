// Correlation cost volume on CUDA cores: one warp per output pixel, vectorised channel
// loads, warp-shuffle (transpose) reduction.  Any C / P / dilation_patch, fp32 or bf16
// storage, fp32 math — the general path and the fp32 parity path of stm_correlation_fwd.
// The bf16 hot path for the STMask shapes is the tcgen05 kernel in corr_tc.cu.
//
//   out[b, ph*P+pw, y, x] = post(scale * sum_c x1[b,y,x,c] * x2[b, y+(ph-r)d, x+(pw-r)d, c])
//
// Replaces correlation_cuda_forward_kernel of spatial_correlation_sampler plus the
// elementwise tail of correlate() and CandidateShift (reference
// track_to_segment_head.py:53-62, TF_utils.py:28-31).
#include "common.cuh"

namespace stm {
namespace {

constexpr int WARPS = 8;

template <typename T> struct Chunk8;  // 8 consecutive channels
template <> struct Chunk8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Chunk8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const uint32_t r[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(r[i] << 16);
      v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
  }
};

template <typename OT>
__device__ __forceinline__ void store_post(const StmCorrDesc& d, OT* out, int64_t idx, float v) {
  v *= d.scale;
  if (d.flags & STM_CORR_LEAKY_RELU) v = v > 0.f ? v : v * d.leaky_slope;
  if (d.flags & STM_CORR_RELU) v = fmaxf(v, 0.f);
  out[idx] = from_f32<OT>(v);
}

// Sum `part[j]` (j = lane-local index of 32 displacements) across the warp so that lane l
// ends up with the total of displacement l: 31 shuffles instead of 160.
__device__ __forceinline__ float transpose_reduce(float (&part)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = upper ? part[i + s] : part[i];
      const float send = upper ? part[i] : part[i + s];
      part[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return part[0];
}

// NCH = number of 256-channel chunks held in registers (C <= 256*NCH, C % 8 == 0); NCH == 0 is
// the scalar fallback for unaligned channel counts.
template <typename T, typename OT, typename FT, int NCH>
__global__ void __launch_bounds__(WARPS * 32)
corr_simt_kernel(const StmCorrDesc d, const T* __restrict__ x1, const T* __restrict__ x2,
                 const FT* __restrict__ fa, const FT* __restrict__ fb, OT* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t pix = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t npix = (int64_t)d.batch * d.h * d.w;
  if (pix >= npix) return;
  const int b = (int)(pix / (d.h * d.w));
  const int rem = (int)(pix - (int64_t)b * d.h * d.w);
  const int y = rem / d.w, x = rem - y * d.w;
  const int P = d.patch, r = P / 2, dl = d.dilation_patch;
  const T* p1 = x1 + b * d.x1_stride_n + y * d.x1_stride_h + x * d.x1_stride_w;
  const T* p2b = x2 + b * d.x2_stride_n;
  const int64_t obase = b * d.out_stride_n + y * d.out_stride_h + x * d.out_stride_w;

  Chunk8<T> a[NCH > 0 ? NCH : 1];
  if (NCH > 0) {
#pragma unroll
    for (int j = 0; j < (NCH > 0 ? NCH : 1); ++j) {
      const int c = j * 256 + lane * 8;
      if (c < d.c) a[j].load(p1 + c);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[j].v[i] = 0.f;
      }
    }
  }

  const int PP = P * P;
  for (int k0 = 0; k0 < PP; k0 += 32) {
    float part[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      part[j] = 0.f;
      const int k = k0 + j;
      if (k < PP) {
        const int ph = k / P, pw = k - ph * P;
        const int y2 = y + (ph - r) * dl, x2c = x + (pw - r) * dl;
        if (y2 >= 0 && y2 < d.h && x2c >= 0 && x2c < d.w) {
          const T* p2 = p2b + y2 * d.x2_stride_h + x2c * d.x2_stride_w;
          float s = 0.f;
          if (NCH > 0) {
#pragma unroll
            for (int jj = 0; jj < (NCH > 0 ? NCH : 1); ++jj) {
              const int c = jj * 256 + lane * 8;
              if (c < d.c) {
                Chunk8<T> bv;
                bv.load(p2 + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fmaf(a[jj].v[i], bv.v[i], s);
              }
            }
          } else {
            for (int c = lane; c < d.c; c += 32) s = fmaf(to_f32(p1[c]), to_f32(p2[c]), s);
          }
          part[j] = s;
        }
      }
    }
    const float tot = transpose_reduce(part, lane);
    const int k = k0 + lane;
    if (k < PP) store_post<OT>(d, out, obase + k * d.out_stride_c, tot);
  }

  if (d.flags & STM_CORR_COPY_FEATS) {
    const bool relu = (d.flags & STM_CORR_RELU) != 0;
    const FT* pa = fa + b * d.feat_a_stride_n + y * d.feat_a_stride_h + x * d.feat_a_stride_w;
    const FT* pb = fb + b * d.feat_b_stride_n + y * d.feat_b_stride_h + x * d.feat_b_stride_w;
    const int foff = d.feat_c_offset > 0 ? d.feat_c_offset : PP;
    for (int c = PP + lane; c < foff; c += 32) out[obase + (int64_t)c * d.out_stride_c] = from_f32<OT>(0.f);   // pad channels
    for (int c = lane; c < d.feat_c; c += 32) {
      float va = to_f32(pa[c]), vb = to_f32(pb[c]);
      if (relu) { va = fmaxf(va, 0.f); vb = fmaxf(vb, 0.f); }
      out[obase + (int64_t)(foff + c) * d.out_stride_c] = from_f32<OT>(va);
      out[obase + (int64_t)(foff + d.feat_c + c) * d.out_stride_c] = from_f32<OT>(vb);
    }
  }
}

template <typename T, typename OT, typename FT>
int launch_typed(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                 cudaStream_t stream) {
  const int64_t npix = (int64_t)d.batch * d.h * d.w;
  if (npix == 0) return STM_OK;
  const dim3 grid((unsigned)((npix + WARPS - 1) / WARPS));
  const bool vec = (d.c % 8 == 0) && (d.x1_stride_w % 8 == 0) && (d.x1_stride_h % 8 == 0) && (d.x1_stride_n % 8 == 0) &&
                   (d.x2_stride_w % 8 == 0) && (d.x2_stride_h % 8 == 0) && (d.x2_stride_n % 8 == 0) &&
                   ((uintptr_t)x1 % 16 == 0) && ((uintptr_t)x2 % 16 == 0) && d.c <= 1024;
  const T* a = (const T*)x1; const T* b = (const T*)x2;
  const FT* f1 = (const FT*)fa; const FT* f2 = (const FT*)fb;
  OT* o = (OT*)out;
  if (!vec) corr_simt_kernel<T, OT, FT, 0><<<grid, WARPS * 32, 0, stream>>>(d, a, b, f1, f2, o);
  else if (d.c <= 256) corr_simt_kernel<T, OT, FT, 1><<<grid, WARPS * 32, 0, stream>>>(d, a, b, f1, f2, o);
  else if (d.c <= 512) corr_simt_kernel<T, OT, FT, 2><<<grid, WARPS * 32, 0, stream>>>(d, a, b, f1, f2, o);
  else corr_simt_kernel<T, OT, FT, 4><<<grid, WARPS * 32, 0, stream>>>(d, a, b, f1, f2, o);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

template <typename T, typename OT>
int launch_feat(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                cudaStream_t s) {
  if ((d.flags & STM_CORR_COPY_FEATS) && d.feat_dtype == STM_F32) return launch_typed<T, OT, float>(d, x1, x2, fa, fb, out, s);
  return launch_typed<T, OT, __nv_bfloat16>(d, x1, x2, fa, fb, out, s);
}

}  // namespace

int launch_corr_simt(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                     cudaStream_t s) {
  if (d.dtype == STM_F32) {
    if (d.out_dtype == STM_F32) return launch_feat<float, float>(d, x1, x2, fa, fb, out, s);
    return launch_feat<float, __nv_bfloat16>(d, x1, x2, fa, fb, out, s);
  }
  if (d.out_dtype == STM_F32) return launch_feat<__nv_bfloat16, float>(d, x1, x2, fa, fb, out, s);
  return launch_feat<__nv_bfloat16, __nv_bfloat16>(d, x1, x2, fa, fb, out, s);
}

}  // namespace stm
