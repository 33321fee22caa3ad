// Microbenchmark: per-SM gather bandwidth of LDG.128 (L1-resident) vs LDS.128 for the DCN access pattern
// (8 lanes share one 128-byte line, 4 lines per warp instruction).   nvcc -arch=sm_100a -O3 gather_bw.cu -o gather_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0 = LDG.128 4 lines/instr, 1 = LDS.128 4 lines/instr, 2 = LDG.32 1 line/instr, 3 = LDG.64 2 lines/instr
__global__ void __launch_bounds__(512, 1) k(const uint4* __restrict__ g, int iters, uint32_t* out, long long* cyc, int lines) {
  extern __shared__ uint4 sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < lines * 8; i += blockDim.x) sm[i] = g[i];
  __syncthreads();
  uint32_t acc = 0;
  uint32_t h = warp * 977u + (lane >> 3) * 131u;
  uint64_t ga[8]; uint32_t sa[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    h = h * 1664525u + 1013904223u;
    uint32_t hh = h;
    if (MODE == 2) hh = __shfl_sync(0xffffffffu, h, 0);
    if (MODE == 3) hh = __shfl_sync(0xffffffffu, h, lane & 16);
    const int line = (hh >> 8) % (lines / 2);
    if (MODE == 0) ga[u] = (uint64_t)(g + line * 8 + (lane & 7));
    if (MODE == 1) sa[u] = (uint32_t)__cvta_generic_to_shared(sm + line * 8 + ((lane & 7) ^ (line & 7)));
    if (MODE == 2) ga[u] = (uint64_t)(reinterpret_cast<const uint32_t*>(g) + line * 32 + lane);
    if (MODE == 3) ga[u] = (uint64_t)(reinterpret_cast<const uint2*>(g) + line * 16 + (lane & 15));
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t sh = (uint32_t)((it * 37) & (lines / 2 - 1)) * 128u;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint32_t x, y, z, w;
      if (MODE == 0) { asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(ga[u] + sh)); acc ^= x ^ y ^ z ^ w; }
      if (MODE == 1) { asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(sa[u] + sh)); acc ^= x ^ y ^ z ^ w; }
      if (MODE == 2) { asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(ga[u] + sh)); acc ^= x; }
      if (MODE == 3) { asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "l"(ga[u] + sh)); acc ^= x ^ y; }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = acc;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const uint4* g, uint32_t* out, long long* cyc, int lines, int bytes_per_instr) {
  const int iters = 20000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, lines * 128);
  k<MODE><<<148, 512, lines * 128>>>(g, 10, out, cyc, lines);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, 512, lines * 128>>>(g, iters, out, cyc, lines);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c[148];
  cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += c[i];
  avg /= 148;
  const double bytes = (double)iters * 8 * 16 * bytes_per_instr;   // 16 warps
  printf("%-28s lines=%4d: %.1f B/clk/SM  (%.2f cyc per warp-instr)  wall %.3f ms -> %.1f GB/s/SM err=%s\n", name, lines, bytes / avg, avg / (iters * 8.0 * 16), ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  uint4* g; uint32_t* out; long long* cyc;
  cudaMalloc(&g, 1 << 20); cudaMemset(g, 1, 1 << 20);
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int lines : {256, 512}) {
    run<0>("LDG.128 4 lines/instr", g, out, cyc, lines, 512);
    run<1>("LDS.128 4 lines/instr", g, out, cyc, lines, 512);
    run<2>("LDG.32  1 line/instr", g, out, cyc, lines, 128);
    run<3>("LDG.64  2 lines/instr", g, out, cyc, lines, 256);
  }
  return 0;
}
