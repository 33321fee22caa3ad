// Layout / elementwise helpers: weight packing (OIHW -> OHWI), NCHW <-> NHWC with dtype
// conversion, and the closed-form FCB(ali) offsets (reference Featurealign.py:46-69).
#include <type_traits>

#include "common.cuh"

namespace stm {
namespace {

template <typename S, typename D>
__global__ void pack_ohwi_kernel(const S* __restrict__ src, D* __restrict__ dst, int O, int I, int K) {
  // src [O][I][K] -> dst [O][K][I]
  const int64_t n = (int64_t)O * I * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const int64_t t = idx / I;
    const int k = (int)(t % K);
    const int o = (int)(t / K);
    dst[idx] = from_f32<D>(to_f32(src[((int64_t)o * I + i) * K + k]));
  }
}

// [n][rows][cols] -> [n][cols][rows] through a 32x33 shared tile
template <typename S, typename D>
__global__ void transpose_kernel(const S* __restrict__ src, D* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int64_t nb = (int64_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = to_f32(src[nb + (int64_t)r * cols + c]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[nb + (int64_t)c * rows + r] = from_f32<D>(tile[threadIdx.x][j]);
  }
}

template <typename S, typename D>
__global__ void ali_offsets_kernel(const S* __restrict__ shape, D* __restrict__ off, int B, int H, int W, int kh, int kw,
                                   int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w,
                                   int64_t o_n, int64_t o_c, int64_t o_h, int64_t o_w) {
  const int64_t n = (int64_t)B * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int64_t t = idx / W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const int64_t sb = b * s_n + y * s_h + x * s_w;
    const float tx = to_f32(shape[sb]), ty = to_f32(shape[sb + s_c]);
    const float tw = to_f32(shape[sb + 2 * s_c]), th = to_f32(shape[sb + 3 * s_c]);
    const float dx = tx * 0.1f * (float)kw, dy = ty * 0.1f * (float)kh;
    const float ew = expf(tw * 0.2f) - 1.0f, eh = expf(th * 0.2f) - 1.0f;
    const int64_t ob = b * o_n + y * o_h + x * o_w;
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j) {
        const int k = i * kw + j;
        off[ob + (int64_t)(2 * k) * o_c] = from_f32<D>(dy + eh * (float)(i - kh / 2));
        off[ob + (int64_t)(2 * k + 1) * o_c] = from_f32<D>(dx + ew * (float)(j - kw / 2));
      }
  }
}

// FCB(ada) offsets: the 1x1 `conv_offset` of Featurealign.py:20-25,44 — a [OC x 4] linear map of the
// box deltas per pixel (no bias).  w is float32 [OC][4].
template <typename S, typename D>
__global__ void ada_offsets_kernel(const S* __restrict__ shape, const float* __restrict__ w, D* __restrict__ off, int B, int H,
                                   int W, int OC, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w,
                                   int64_t o_n, int64_t o_c, int64_t o_h, int64_t o_w) {
  extern __shared__ float ws[];   // OC * 4
  for (int i = threadIdx.x; i < OC * 4; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int64_t n = (int64_t)B * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int64_t t = idx / W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const int64_t sb = b * s_n + y * s_h + x * s_w;
    const float t0 = to_f32(shape[sb]), t1 = to_f32(shape[sb + s_c]);
    const float t2 = to_f32(shape[sb + 2 * s_c]), t3 = to_f32(shape[sb + 3 * s_c]);
    const int64_t ob = b * o_n + y * o_h + x * o_w;
    for (int o = 0; o < OC; ++o) {
      const float v = fmaf(ws[4 * o + 3], t3, fmaf(ws[4 * o + 2], t2, fmaf(ws[4 * o + 1], t1, ws[4 * o] * t0)));
      off[ob + (int64_t)o * o_c] = from_f32<D>(v);
    }
  }
}

template <typename F>
int dispatch2(int sd, int dd, F&& f) {
  if (sd == STM_F32 && dd == STM_F32) return f((const float*)nullptr, (float*)nullptr);
  if (sd == STM_F32 && dd == STM_BF16) return f((const float*)nullptr, (__nv_bfloat16*)nullptr);
  if (sd == STM_BF16 && dd == STM_F32) return f((const __nv_bfloat16*)nullptr, (float*)nullptr);
  if (sd == STM_BF16 && dd == STM_BF16) return f((const __nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr);
  set_error("unknown dtype %d/%d", sd, dd);
  return STM_ERR_INVALID_ARGUMENT;
}

}  // namespace

int pack_weight_ohwi(const void* src, int sd, void* dst, int dd, int O, int I, int K, cudaStream_t stream) {
  const int64_t n = (int64_t)O * I * K;
  if (n == 0) return STM_OK;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  return dispatch2(sd, dd, [&](auto* s, auto* d) -> int {
    using S = std::remove_cv_t<std::remove_pointer_t<decltype(s)>>;
    using D = std::remove_pointer_t<decltype(d)>;
    pack_ohwi_kernel<S, D><<<blocks, 256, 0, stream>>>((const S*)src, (D*)dst, O, I, K);
    count_launch();
    STM_CUDA_OK(cudaGetLastError());
    return (int)STM_OK;
  });
}

int transpose_batched(const void* src, int sd, void* dst, int dd, int n, int rows, int cols, cudaStream_t stream) {
  if ((int64_t)n * rows * cols == 0) return STM_OK;
  if (n > 65535) { set_error("batch %d too large for the layout kernel", n); return STM_ERR_UNSUPPORTED; }
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, n), block(32, 8);
  if (grid.y > 65535) { set_error("tensor too large for the layout kernel"); return STM_ERR_UNSUPPORTED; }
  return dispatch2(sd, dd, [&](auto* s, auto* d) -> int {
    using S = std::remove_cv_t<std::remove_pointer_t<decltype(s)>>;
    using D = std::remove_pointer_t<decltype(d)>;
    transpose_kernel<S, D><<<grid, block, 0, stream>>>((const S*)src, (D*)dst, rows, cols);
    count_launch();
    STM_CUDA_OK(cudaGetLastError());
    return (int)STM_OK;
  });
}

int ali_offsets(const void* shape, const int64_t* ss, int sd, void* off, const int64_t* os, int od, int B, int H, int W,
                int kh, int kw, cudaStream_t stream) {
  const int64_t n = (int64_t)B * H * W;
  if (n == 0) return STM_OK;
  const int blocks = (int)((n + 127) / 128 < 8192 ? (n + 127) / 128 : 8192);
  return dispatch2(sd, od, [&](auto* s, auto* d) -> int {
    using S = std::remove_cv_t<std::remove_pointer_t<decltype(s)>>;
    using D = std::remove_pointer_t<decltype(d)>;
    ali_offsets_kernel<S, D><<<blocks, 128, 0, stream>>>((const S*)shape, (D*)off, B, H, W, kh, kw, ss[0], ss[1], ss[2],
                                                         ss[3], os[0], os[1], os[2], os[3]);
    count_launch();
    STM_CUDA_OK(cudaGetLastError());
    return (int)STM_OK;
  });
}

int ada_offsets(const void* shape, const int64_t* ss, int sd, const float* w, void* off, const int64_t* os, int od, int B,
                int H, int W, int OC, cudaStream_t stream) {
  const int64_t n = (int64_t)B * H * W;
  if (n == 0) return STM_OK;
  const int blocks = (int)((n + 127) / 128 < 8192 ? (n + 127) / 128 : 8192);
  return dispatch2(sd, od, [&](auto* s, auto* d) -> int {
    using S = std::remove_cv_t<std::remove_pointer_t<decltype(s)>>;
    using D = std::remove_pointer_t<decltype(d)>;
    ada_offsets_kernel<S, D><<<blocks, 128, OC * 4 * sizeof(float), stream>>>(
        (const S*)shape, w, (D*)off, B, H, W, OC, ss[0], ss[1], ss[2], ss[3], os[0], os[1], os[2], os[3]);
    count_launch();
    STM_CUDA_OK(cudaGetLastError());
    return (int)STM_OK;
  });
}

}  // namespace stm
