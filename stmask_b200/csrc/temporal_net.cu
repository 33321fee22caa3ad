// Tail of TemporalNet (reference layers/modules/track_to_segment_head.py:17-19,31-35): AvgPool2d(7x7) over the
// 7x7x1024 conv3 output of every box, then the two Linear layers (fc: 1024 -> 4 box deltas, fc_coeff: 1024 -> 32
// mask coefficients) — one kernel, one CTA per box.  The three 3x3 convs in front of it run on the tcgen05 main loop
// of dcn_tc.cu in its plain-conv mode.
//
//   x [n, hw, c]  NHWC activations (bf16 or fp32), hw = 49, c = 1024
//   w [out, c]    fp32, rows of fc followed by rows of fc_coeff;  b [out] fp32
//   y [n, out]    fp32:  y = W * mean_hw(x) + b
#include "common.cuh"

namespace stm {
namespace {

constexpr int PF_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(PF_THREADS) pool_fc_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, float* __restrict__ y, int hw, int c,
                                                             int out_features, int64_t x_stride_n, int64_t x_stride_p) {
  extern __shared__ float pooled[];                       // [c]
  const int n = blockIdx.x;
  const T* xn = x + (int64_t)n * x_stride_n;
  const float inv = 1.f / (float)hw;
  // pooling: threads over channels (coalesced along c), loop over the pixels; fp32 accumulation
  for (int ch = threadIdx.x; ch < c; ch += PF_THREADS) {
    float acc = 0.f;
    for (int p = 0; p < hw; ++p) acc += to_f32(xn[(int64_t)p * x_stride_p + ch]);
    pooled[ch] = acc * inv;
  }
  __syncthreads();
  // the two linear layers: one warp per output feature, shuffle reduction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < out_features; o += PF_THREADS / 32) {
    const float* wr = w + (int64_t)o * c;
    float acc = 0.f;
    for (int ch = lane; ch < c; ch += 32) acc = fmaf(__ldg(wr + ch), pooled[ch], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) y[(int64_t)n * out_features + o] = acc + (b != nullptr ? __ldg(b + o) : 0.f);
  }
}

}  // namespace

int launch_pool_fc(const void* x, int dtype, int n, int hw, int c, int64_t x_stride_n, int64_t x_stride_p, const float* w,
                   const float* b, int out_features, float* y, cudaStream_t stream) {
  const size_t smem = (size_t)c * sizeof(float);
  if (smem > 48 * 1024) { set_error("pool_fc: %d channels do not fit the pooled vector in shared memory", c); return STM_ERR_UNSUPPORTED; }
  if (dtype == STM_BF16)
    pool_fc_kernel<__nv_bfloat16><<<n, PF_THREADS, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, b, y, hw, c, out_features,
                                                                 x_stride_n, x_stride_p);
  else
    pool_fc_kernel<float><<<n, PF_THREADS, smem, stream>>>(reinterpret_cast<const float*>(x), w, b, y, hw, c, out_features, x_stride_n,
                                                         x_stride_p);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace stm
