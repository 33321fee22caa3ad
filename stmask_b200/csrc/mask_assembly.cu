// Mask assembly and mask IoU for a batch of frames, on the device (the reference does it per frame in torch:
// generate_mask, layers/mask_utils.py:111-128 with crop, layers/box_utils.py:341-364, and mask_iou,
// layers/box_utils.py:435-447, called from Track_TF.track, layers/functions/track_TF.py:76-112).
//
//   mask[f, n, y, x] = crop_n(sigmoid( proto[f, y, x, :] . tanh(coeff[f, n, :]) ))
//   crop_n keeps the pixels with x1 <= x < x2 and y1 <= y < y2, where (x1, x2) = sanitize(box.x1 * w, box.x2 * w) with
//   1 pixel of padding, clamped to [0, w] (box_utils.py:298-317), same for y.
// The kernel also writes the masks thresholded at 0.5 as BIT planes (one 32-bit word per 32 pixels), on which
// mask_iou_kernel computes intersection / union with popcounts instead of an [n1, hw] x [hw, n2] float GEMM.
#include "common.cuh"

namespace stm {
namespace {

constexpr int MA_THREADS = 256;
constexpr int MA_MAX_K = 64;

__global__ void __launch_bounds__(MA_THREADS) mask_assembly_kernel(const float* __restrict__ proto, const float* __restrict__ coeff,
                                                                   const float* __restrict__ boxes, const int32_t* __restrict__ count,
                                                                   float* __restrict__ masks, uint32_t* __restrict__ bits, int h, int w,
                                                                   int k, int max_n, int words) {
  __shared__ float s_c[MA_MAX_K];
  const int f = blockIdx.z, n = blockIdx.y;
  const int n_valid = count ? min(count[f], max_n) : max_n;
  if (n >= n_valid) return;
  const float* cf = coeff + ((size_t)f * max_n + n) * k;
  if (threadIdx.x < k) s_c[threadIdx.x] = tanhf(cf[threadIdx.x]);          // cfg.mask_proto_coeff_activation
  __syncthreads();
  const float* b = boxes + ((size_t)f * max_n + n) * 4;
  // sanitize_coordinates(_x1, _x2, img_size, padding = 1, cast = False)
  const float bx1 = b[0] * (float)w, bx2 = b[2] * (float)w, by1 = b[1] * (float)h, by2 = b[3] * (float)h;
  const float x1 = fmaxf(fminf(bx1, bx2) - 1.f, 0.f), x2 = fminf(fmaxf(bx1, bx2) + 1.f, (float)w);
  const float y1 = fmaxf(fminf(by1, by2) - 1.f, 0.f), y2 = fminf(fmaxf(by1, by2) + 1.f, (float)h);
  const int hw = h * w;
  const int pix = blockIdx.x * MA_THREADS + threadIdx.x;
  float v = 0.f;
  if (pix < hw) {
    const int y = pix / w, x = pix - y * w;
    if ((float)x >= x1 && (float)x < x2 && (float)y >= y1 && (float)y < y2) {
      const float* pr = proto + ((size_t)f * hw + pix) * k;
      float acc = 0.f;
      for (int c = 0; c < k; ++c) acc = fmaf(pr[c], s_c[c], acc);
      v = 1.f / (1.f + expf(-acc));                                        // cfg.mask_proto_mask_activation
    }
    masks[((size_t)f * max_n + n) * hw + pix] = v;
  }
  const uint32_t word = __ballot_sync(0xffffffffu, v > 0.5f);
  if ((threadIdx.x & 31) == 0 && (pix >> 5) < words) bits[((size_t)f * max_n + n) * words + (pix >> 5)] = word;
}

// iou[f, i, j] = |A_i & B_j| / |A_i | B_j|  (0 when the union is empty), A = bit masks of frame f in set 1, B in set 2
__global__ void __launch_bounds__(256) mask_iou_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                       const int32_t* __restrict__ na, const int32_t* __restrict__ nb, float* __restrict__ iou,
                                                       int max_a, int max_b, int words) {
  const int f = blockIdx.z;
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  const int va = na ? min(na[f], max_a) : max_a, vb = nb ? min(nb[f], max_b) : max_b;
  if (i >= va || j >= vb) return;
  const uint32_t* pa = a + ((size_t)f * max_a + i) * words;
  const uint32_t* pb = b + ((size_t)f * max_b + j) * words;
  int inter = 0, uni = 0;
  for (int t = 0; t < words; ++t) {
    const uint32_t x = __ldg(pa + t), y = __ldg(pb + t);
    inter += __popc(x & y);
    uni += __popc(x | y);
  }
  iou[((size_t)f * max_a + i) * max_b + j] = uni > 0 ? (float)inter / (float)uni : 0.f;
}

}  // namespace

int launch_mask_assembly(const float* proto, const float* coeff, const float* boxes, const int32_t* count, float* masks,
                         uint32_t* bits, int frames, int h, int w, int k, int max_n, cudaStream_t stream) {
  const int hw = h * w, words = (hw + 31) / 32;
  dim3 grid((hw + MA_THREADS - 1) / MA_THREADS, max_n, frames);
  mask_assembly_kernel<<<grid, MA_THREADS, 0, stream>>>(proto, coeff, boxes, count, masks, bits, h, w, k, max_n, words);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

int launch_mask_iou(const uint32_t* a, const uint32_t* b, const int32_t* na, const int32_t* nb, float* iou, int frames, int max_a,
                    int max_b, int words, cudaStream_t stream) {
  dim3 grid((max_b + 255) / 256, max_a, frames);
  mask_iou_kernel<<<grid, 256, 0, stream>>>(a, b, na, nb, iou, max_a, max_b, words);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace stm
