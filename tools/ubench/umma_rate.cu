// How fast does one CTA retire tcgen05.mma (M = 128, K = 16, bf16, SS mode) as a function of N, of the A operand's start
// alignment (1024-byte aligned vs shifted by rows of 128 bytes, as conv_tma.cu's tap views are) and of how many independent
// TMEM accumulators the stream of MMAs rotates over?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I stmask_b200/csrc tools/ubench/umma_rate.cu -o tools/ubench/umma_rate
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace stm::tc;

__global__ void __launch_bounds__(256, 1) k(long long* out, int n, int shift_rows, int accs, int iters, int mstep, int spin) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;                        // 512 rows x 128 B
  uint8_t* sb = smem + 512 * 128;            // 256 rows x 128 B
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sb + 256 * 128);
  uint64_t* sbar = mbar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 2);
  for (int i = threadIdx.x; i < (512 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); mbar_init(sbar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // one "tap": 4 MMAs (K = 64) on a view shifted by (it % mstep) * shift_rows rows
      const uint32_t a0 = smem_u32(sa) + (uint32_t)((it % mstep) * shift_rows) * 128u;
      const uint64_t ad = umma_desc_sw128(a0), bd = umma_desc_sw128(smem_u32(sb));
      const uint32_t d = tm + (uint32_t)((it % accs) * n);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_bf16(d, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, 1u);
    }
    umma_commit(mbar);
    mbar_wait(mbar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    mbar_arrive(sbar);
  } else if (warp >= 4 && warp < 4 + spin) {
    if (spin >= 10) mbar_wait_relaxed(sbar, 0); else mbar_wait(sbar, 0);      // spinning bystanders, like an epilogue waiting for its accumulator
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 1024);
  const int smem = (512 + 256) * 128 + 64 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int n : {32, 256})
    for (int accs : {1})
      for (int shift : {43})
       for (int grid : {1, 2, 74, 148, 296}) { const int spin = 0;
        if (accs * n > 512) continue;
        long long h = 0, hs[296];
        for (int rep = 0; rep < 2; ++rep) {
          k<<<grid, 256, smem>>>(d, n, shift, accs, iters, 3, spin);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          cudaMemcpy(hs, d, 8 * grid, cudaMemcpyDeviceToHost);
          h = 0; for (int i = 0; i < grid; ++i) h = hs[i] > h ? hs[i] : h;
        }
        printf("grid %3d: ", grid);
        printf("N=%3d accumulators=%d A shift step=%2d rows, %d spinning warps%s: %.1f clk per MMA (K=16)\n", n, accs, shift, spin % 10, spin >= 10 ? " (nanosleep backoff)" : "", (double)h / (iters * 4));
      }
  return 0;
}
