// Experiment: can a tcgen05 SS-MMA read its K-major, 128B-swizzled A operand from a start address that is shifted by
// p ROWS (p * 128 bytes, not 1024-aligned)?  This is what a regular convolution needs to reuse ONE haloed input tile
// for all taps (flattened padded image: tap (i, j) = the same tile shifted by i * Wh + j pixels).
//   variant 0: descriptor base_offset = 0            variant 1: base_offset = (start >> 7) & 7
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I stmask_b200/csrc tools/ubench/umma_shift.cu -o tools/ubench/umma_shift -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace stm::tc;

constexpr int ROWS = 256, NB = 32;   // A tile: 256 pixel rows x 64 ch; B: 32 x 64

__device__ __forceinline__ uint64_t desc_sw128_off(uint32_t addr, uint32_t base_off) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, float* out,
                                            int shift, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;                       // 256 x 128 B
  uint8_t* sb = smem + ROWS * 128;          // 32 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + NB * 128);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 32); tmem_relinquish(); }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, ROWS * 128 + NB * 128);
    tma_load_2d(sa, &ma, bar, 0, 0);
    tma_load_2d(sb, &mb, bar, 0, 0);
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const uint32_t a0 = smem_u32(sa) + (uint32_t)shift * 128u;
    const uint32_t bo = variant ? (a0 >> 7) & 7u : 0u;
    const uint32_t idesc = umma_idesc_bf16(128, NB);
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(tm, desc_sw128_off(a0, bo) + (uint64_t)(2 * kk), umma_desc_sw128(smem_u32(sb)) + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
    umma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tcgen05_fence_after();
  uint32_t acc[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), acc);
  tmem_ld_wait();
  for (int i = 0; i < NB; ++i) out[(warp * 32 + lane) * NB + i] = __uint_as_float(acc[i]);
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 32); }
}

int main() {
  std::vector<__nv_bfloat16> ha(ROWS * 64), hb(NB * 64);
  srand(1);
  for (auto& v : ha) v = __float2bfloat16((rand() % 17 - 8) / 8.f);
  for (auto& v : hb) v = __float2bfloat16((rand() % 9 - 4) / 4.f);
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * NB * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  cuuint64_t dA[2] = {64, ROWS}, dB[2] = {64, NB}; cuuint64_t st[1] = {128}; cuuint32_t bA[2] = {64, ROWS}, bB[2] = {64, NB}, es[2] = {1, 1};
  cuInit(0);
  if (cuTensorMapEncodeTiled(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dA, st, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
      cuTensorMapEncodeTiled(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dB, st, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 1; }
  const int smem = ROWS * 128 + NB * 128 + 64 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(128 * NB);
  for (int variant = 0; variant < 2; ++variant)
    for (int shift : {0, 1, 2, 3, 5, 8, 11, 42, 83}) {
      k<<<1, 128, smem>>>(ma, mb, dout, shift, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d shift %d: %s\n", variant, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NB; ++n) {
          double ref = 0;
          for (int c = 0; c < 64; ++c) ref += (double)__bfloat162float(ha[(m + shift) * 64 + c]) * __bfloat162float(hb[n * 64 + c]);
          worst = fmax(worst, fabs(ref - ho[m * NB + n]));
        }
      printf("variant %d (base_offset %s) shift %2d rows: max |err| = %g %s\n", variant, variant ? "(start>>7)&7" : "0", shift, worst, worst < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
