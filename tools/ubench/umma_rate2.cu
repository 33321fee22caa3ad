// conv_tma.cu's MMA stream replayed without TMA / epilogue: 3 A stages of 27 KB, 36 resident weight slices of 4 KB (N = 32),
// 9 taps per chunk as shifted views, 4 MMAs per tap.  Which ingredient costs the 2x over umma_rate.cu's 77 clk per MMA?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I stmask_b200/csrc tools/ubench/umma_rate2.cu -o tools/ubench/umma_rate2
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace stm::tc;

// whole-warp form: every lane executes the instruction stream, one elected lane issues (predicate inside the asm block)
__device__ __forceinline__ void umma_bf16_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
constexpr int A_STAGE = 27648, NSTAGE = 3, B_SLOT = 4096, NB = 36;

__global__ void __launch_bounds__(128, 1) k(long long* out, int n, int mode, int tiles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;
  uint8_t* sb = smem + NSTAGE * A_STAGE;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sb + NB * B_SLOT);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 8);
  for (int i = threadIdx.x; i < (NSTAGE * A_STAGE + NB * B_SLOT) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(mbar + i, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = *slot;
  if (mode & 16) {
   if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const long long t0 = clock64();
    uint32_t ac = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tm + (uint32_t)((t & 1) * n);
      for (int c = 0; c < 4; ++c, ++ac) {
        const uint32_t a_base = smem_u32(sa) + (ac % NSTAGE) * A_STAGE;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int i = tap / 3, j = tap - 3 * i;
          const uint32_t a_addr = a_base + (uint32_t)((i * 42 + j) * 128);
          const uint32_t b_addr = smem_u32(sb) + (uint32_t)((c * 9 + tap) * B_SLOT);
          const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_warp(d, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, (c | tap | kk) != 0 ? 1u : 0u);
        }
      }
    }
    if (threadIdx.x == 0) { umma_commit(mbar); mbar_wait(mbar, 0); out[blockIdx.x] = clock64() - t0; }
    __syncwarp();
   }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const long long t0 = clock64();
    uint32_t ac = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tm + (uint32_t)((t & 1) * n);
      for (int c = 0; c < 4; ++c, ++ac) {
        const uint32_t a_base = smem_u32(sa) + ((mode & 1) ? (ac % NSTAGE) * A_STAGE : 0u);
        for (int tap = 0; tap < 9; ++tap) {
          const int i = tap / 3, j = tap - 3 * i;
          const uint32_t a_addr = a_base + ((mode & 2) ? (uint32_t)((i * 42 + j) * 128) : 0u);
          const uint32_t b_addr = smem_u32(sb) + ((mode & 4) ? (uint32_t)((c * 9 + tap) * B_SLOT) : 0u);
          const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, (c | tap | kk) != 0 ? 1u : 0u);
        }
        if (mode & 8) umma_commit(mbar + 1 + (ac % NSTAGE));
      }
      if (mode & 8) umma_commit(mbar + 4 + (t & 1));
    }
    umma_commit(mbar);
    mbar_wait(mbar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 1024);
  const int smem = NSTAGE * A_STAGE + NB * B_SLOT + 128 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = 56;
  for (int n : {32, 64, 128, 256})
    for (int mode : {15, 16}) {
      if (n > 32) continue;      // the weight slices are sized for N = 32
      long long hs[148], h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        k<<<148, 128, smem>>>(d, n, mode, tiles);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        cudaMemcpy(hs, d, 8 * 148, cudaMemcpyDeviceToHost);
        h = 0; for (int i = 0; i < 148; ++i) h = hs[i] > h ? hs[i] : h;
      }
      printf("N=%3d mode %2d (1: rotate A stages, 2: shifted tap views, 4: weight slice per tap, 8: commits): %.1f clk per MMA\n", n, mode,
             (double)h / (tiles * 144));
    }
  return 0;
}
