// Regular (zero-offset) convolution as an implicit GEMM whose A operand never passes through registers.
//
//   Y[M, N] = A[M, K] * W[N, K]^T      M = output pixels of a tile, N = Cout, K = kh*kw*Cin
//
// The deformable kernel (dcn_tc.cu) has to build every A row with a gather; a regular convolution does not: tap (i, j)
// of a tile is the SAME haloed input tile shifted by i rows and j columns.  So per 64-channel chunk ONE 4-D TMA box
// [images][BH = TH + kh - 1][BW = TW + kw - 1][64 ch] lands in shared memory (128-byte swizzle, out-of-map pixels
// zero-filled by TMA = the convolution's padding) and every tap's A operand is a UMMA descriptor whose start address is
// that tile shifted by (i * BW + j) rows of 128 bytes — the swizzle is a function of the shared-memory ADDRESS, so a
// start that is not 1024-byte aligned reads the right chunks (tools/ubench/umma_shift.cu checks exactly this on a B200).
// GEMM row m = r * BW + c is output pixel (r, c) of the tile; the kw - 1 columns per row that straddle the halo are
// computed and dropped (TH x TW is chosen per feature map to keep >= 94 % of the 128 rows live on the FPN / backbone maps).
// Stride 2 (the offset predictors of the stride-2 bottlenecks): the four parity sub-lattices of the input are four
// such tiles, loaded with TMA element strides {1, 2, 2, 1}; every tap reads one of them, again as a shifted view.
//
// Persistent, one CTA per SM, warp-specialised:
//   warp 0     TMA: input tiles (ring of `sa` stages) and [N x 64] weight slices (ring of `sb` slots; when all
//              kh*kw*Cin/64 slices fit they are loaded ONCE per CTA and stay resident)
//   warp 1     MMA: tcgen05.mma M = 128, N <= 256, bf16 -> fp32 into one of TWO TMEM accumulators
//   warps 2-5  epilogue of the previous tile while the next one is multiplied: tcgen05.ld, + bias, ReLU,
//              bf16 or fp32 (STM_DCN_OUT_F32) NHWC stores
//
// Used for the DCN offset / mask predictors (backbone.py:24-26), the prediction-head convs
// (prediction_head_FC.py:71-127,150-183) and, when their crops tile well, the TemporalNet convs
// (track_to_segment_head.py:14-16).  Reached through stm_deform_conv2d_fwd with STM_DCN_ZERO_OFFSET.
#include <cstdio>

#include "common.cuh"
#include "tc_common.cuh"

namespace stm {
namespace {

using namespace tc;

constexpr int MAX_SA = 6;
constexpr int MAX_SB = 40;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int MAX_TAPS = 25;

struct ConvTmaProb {
  void* y;
  int64_t y_sn, y_sh, y_sw, y_sc;      // y_sc: channel stride (1 = NHWC; out_h * y_sh for STM_DCN_OUT_PLANAR)
  int32_t batch, out_h, out_w;
  int32_t tw, th, bb, bw, bh;          // tile: output columns / rows / images; TMA box columns / rows (per parity tile)
  int32_t tiles_x, tiles_y, tiles_b;
  int32_t tile_begin;                  // first spatial tile of this problem in the launch
  int32_t box_bytes;                   // bytes one TMA box writes (per parity tile)
};

struct ConvTmaArgs {
  ConvTmaProb prob[STM_DCN_MAX_PROBLEMS];
  int32_t n_probs, total_tiles, n_tiles, block_n;
  int32_t chunks, kh, kw, ph, pw;
  int32_t sa, sb, a_stage_bytes, a_phase_bytes, resident;
  int32_t tmem_cols, flags;
  int32_t exp_;                        // experiment bits (profiling only)
  int32_t fuse_kw;                     // > 0: the kw horizontal taps are fused into the GEMM's N (see FUSE below)
  int32_t n_mma, pad_;                 // UMMA N = block_n, or block_n * kw when fused
  const float* bias;
};

struct ConvTmaMaps {
  CUtensorMap x[STM_DCN_MAX_PROBLEMS];
  CUtensorMap w;
};

struct TileAt {
  int pi, nt, b0, h0, w0;
};

__device__ __forceinline__ TileAt decode_tile(const ConvTmaArgs& a, int tile) {
  TileAt t;
  const int sp = tile / a.n_tiles;
  t.nt = tile - sp * a.n_tiles;
  t.pi = 0;
#pragma unroll 1
  for (int i = 1; i < a.n_probs; ++i)
    if (sp >= a.prob[i].tile_begin) t.pi = i;
  const ConvTmaProb& q = a.prob[t.pi];
  const int local = sp - q.tile_begin;
  const int tx = local % q.tiles_x;
  const int rest = local / q.tiles_x;
  const int ty = rest % q.tiles_y;
  t.b0 = (rest / q.tiles_y) * q.bb;
  t.h0 = ty * q.th;
  t.w0 = tx * q.tw;
  return t;
}

// tap index along one axis -> (parity tile, shift inside it); S = 1: no parities, shift = tap index
template <int S>
__device__ __forceinline__ void tap_axis(int j, int pad, int& parity, int& shift) {
  if (S == 1) { parity = 0; shift = j; return; }
  const int t = j - pad;
  parity = t & 1;
  const int lat0 = -((pad + 1) >> 1);              // floor(-pad / 2)
  shift = ((t - parity) >> 1) - lat0;
}

// FUSE (stride 1, small Cout — the offset / mask predictors): with N = 32 the tensor core spends its time re-reading the A
// tile from shared memory once per tap (profiles/r02_tma_predictor_bb256.txt: operand reads 60 % of the shared-memory pipe,
// tensor pipe 25 %).  The kw horizontal taps are therefore moved from K into N: one MMA per (vertical tap, K slice) with
// B = [Cout x kw] weight rows (a 4-D TMA box of the OHWI weight: row n * kw + j) computes P_j[m] = sum_c x[m] w[n, i, j, c]
// for every input pixel m of the tile, and out[m] = P_0[m] + P_1[m + 1] + P_2[m + 2] is formed in the epilogue with two
// warp shuffles per channel (rows m + 1, m + 2 of the next lane quarter come through shared memory).  A third of the MMAs,
// a third of the A reads.
template <int S, int TAPS, bool FUSE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tma_kernel(const __grid_constant__ ConvTmaArgs a, const __grid_constant__ ConvTmaMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_slot_bytes = a.n_mma * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a.sa * a.a_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + a.sb * b_slot_bytes);
  uint64_t* a_empty = a_full + MAX_SA;
  uint64_t* b_full = a_empty + MAX_SA;
  uint64_t* b_empty = b_full + MAX_SB;
  uint64_t* acc_full = b_empty + MAX_SB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);       // [block_n]
  uint32_t* s_tapoff = reinterpret_cast<uint32_t*>(s_bias + a.block_n);   // [problem][tap]: descriptor offset of the tap's view
  float* s_xchg = reinterpret_cast<float*>(s_tapoff + STM_DCN_MAX_PROBLEMS * MAX_TAPS);   // FUSE: [2][4 quarters][3][block_n]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int taps = FUSE ? a.kh : a.kh * a.kw;                   // A views (and weight slices) per 64-channel chunk
  const int items = a.chunks * taps;                            // weight slices per tile
  const bool resident = a.resident != 0;

  if (tid == 0) {
    for (int i = 0; i < a.n_probs; ++i) prefetch_tensormap(&maps.x[i]);
    prefetch_tensormap(&maps.w);
    for (int s = 0; s < a.sa; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < a.sb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_THREADS / 32); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols); tmem_relinquish(); }
  if (warp >= 2) {
    for (int i = tid - 64; i < a.n_probs * MAX_TAPS; i += EPI_THREADS) {
      const int pi = i / MAX_TAPS, tap = i - pi * MAX_TAPS;
      uint32_t off = 0;
      if (tap < taps) {
        const int ti = FUSE ? tap : tap / a.kw, tj = FUSE ? 0 : tap - ti * a.kw;
        int pyi, sy, pxi, sx;
        tap_axis<S>(ti, a.ph, pyi, sy);
        tap_axis<S>(tj, a.pw, pxi, sx);
        off = (uint32_t)((pyi * 2 + pxi) * a.a_phase_bytes + (sy * a.prob[pi].bw + sx) * 128) >> 4;
      }
      s_tapoff[i] = off;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA ===============================
    if (lane == 0) {
      uint32_t ac = 0, astage = 0, aphase = 0, bslot = 0, bphase = 0;
      bool first = true;
#pragma unroll 1
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const TileAt t = decode_tile(a, tile);
        const ConvTmaProb& q = a.prob[t.pi];
        const CUtensorMap* mx = &maps.x[t.pi];
        const int n0 = t.nt * a.block_n;
#pragma unroll 1
        for (int c = 0; c < a.chunks; ++c) {
          const uint32_t s = astage;
          if (!((a.exp_ & 1) && ac >= (uint32_t)a.sa)) {
          if (a.exp_ & 32) mbar_wait_relaxed(&a_empty[s], aphase ^ 1u); else mbar_wait(&a_empty[s], aphase ^ 1u);
          uint8_t* dst = sA + s * a.a_stage_bytes;
          if (S == 1) {
            mbar_arrive_expect_tx(&a_full[s], (uint32_t)q.box_bytes);
            tma_load_4d(dst, mx, &a_full[s], c * 64, t.w0 - a.pw, t.h0 - a.ph, t.b0);
          } else {
            mbar_arrive_expect_tx(&a_full[s], (uint32_t)q.box_bytes * 4u);
            const int lx = -((a.pw + 1) >> 1), ly = -((a.ph + 1) >> 1);
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
              for (int px = 0; px < 2; ++px)
                tma_load_4d(dst + (py * 2 + px) * a.a_phase_bytes, mx, &a_full[s], c * 64, 2 * (t.w0 + lx) + px, 2 * (t.h0 + ly) + py, t.b0);
          }
          }
          ++ac;
          if (++astage == (uint32_t)a.sa) { astage = 0; aphase ^= 1u; }
          if (resident && !first) continue;
#pragma unroll 1
          for (int tap = 0; tap < taps; ++tap) {
            const int kcol = (tap * a.chunks + c) * 64;         // OHWI: K index = tap * Cin + channel
            if (resident) {
              if (first) {
                const int slot = c * taps + tap;
                mbar_arrive_expect_tx(&b_full[slot], (uint32_t)b_slot_bytes);
                if (FUSE) tma_load_4d(sB + slot * b_slot_bytes, &maps.w, &b_full[slot], c * 64, 0, tap, n0);
                else tma_load_2d(sB + slot * b_slot_bytes, &maps.w, &b_full[slot], kcol, n0);
              }
            } else {
              const uint32_t slot = bslot;
              mbar_wait_relaxed(&b_empty[slot], bphase ^ 1u);
              mbar_arrive_expect_tx(&b_full[slot], (uint32_t)b_slot_bytes);
              if (FUSE) tma_load_4d(sB + slot * b_slot_bytes, &maps.w, &b_full[slot], c * 64, 0, tap, n0);
              else tma_load_2d(sB + slot * b_slot_bytes, &maps.w, &b_full[slot], kcol, n0);
              if (++bslot == (uint32_t)a.sb) { bslot = 0; bphase ^= 1u; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA ===============================
    // ONE thread issues every MMA, and what bounds an N <= 64 convolution is how fast it can do that (tools/ubench/
    // umma_rate2.cu: the tensor pipe retires an M = 128, K = 16 MMA every ~55 clk for N <= 64, a lone thread needs 5-7 clk
    // per dependent instruction): the per-tap work is therefore a table lookup (descriptor offset of the tap's shifted
    // view, 16-byte units), two 64-bit adds and the four MMAs, fully unrolled for 3x3.
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)a.n_mma);
      const uint32_t bstep = (uint32_t)(b_slot_bytes >> 4);
      const uint64_t bdesc0 = umma_desc_sw128(smem_u32(sB));
      uint32_t ac = 0, tl = 0, bslot = 0, bphase = 0, astage = 0, aphase = 0;
      constexpr int NT = TAPS > 0 ? TAPS : 1;
      uint32_t toff[NT] = {};
#ifdef STM_CONV_TMA_TRACE
      long long t_acc = 0, t_a = 0, t_b = 0, t_all = clock64(), tq;
#define TRACE_T0 tq = clock64()
#define TRACE_ADD(x) x += clock64() - tq
#else
#define TRACE_T0
#define TRACE_ADD(x)
#endif
#pragma unroll 1
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tl) {
        const TileAt t = decode_tile(a, tile);
        const uint32_t* tab = s_tapoff + t.pi * MAX_TAPS;
        if (TAPS > 0) {                                                // (a handful of LDS per tile)
#pragma unroll
          for (int i = 0; i < NT; ++i) toff[i] = tab[i];
        }
        const uint32_t buf = tl & 1u;
        TRACE_T0;
        if (a.exp_ & 32) mbar_wait_relaxed(&acc_empty[buf], ((tl >> 1) & 1u) ^ 1u); else mbar_wait(&acc_empty[buf], ((tl >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
        TRACE_ADD(t_acc);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)a.n_mma;
        const bool b_ready = resident && tl > 0 && !(a.exp_ & 8);                       // every weight slice is already in shared memory
#pragma unroll 1
        for (int c = 0; c < a.chunks; ++c) {
          const uint32_t s = astage;
          TRACE_T0;
          if (!((a.exp_ & 1) && ac >= (uint32_t)a.sa)) { if (a.exp_ & 32) mbar_wait_relaxed(&a_full[s], aphase); else mbar_wait(&a_full[s], aphase); }
          TRACE_ADD(t_a);
          tcgen05_fence_after();
          const uint64_t adesc_s = umma_desc_sw128(smem_u32(sA + s * a.a_stage_bytes));
          if (b_ready) {
            uint64_t bd = bdesc0 + (uint64_t)((uint32_t)(c * taps) * bstep);
            if (TAPS > 0) {
#pragma unroll
              for (int tap = 0; tap < NT; ++tap) {
                const uint64_t ad = adesc_s + (uint64_t)toff[tap];
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (c | tap | k) != 0 ? 1u : 0u);
                bd += bstep;
              }
            } else {
#pragma unroll 1
              for (int tap = 0; tap < taps; ++tap) {
                const uint64_t ad = adesc_s + (uint64_t)tab[tap];
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (c | tap | k) != 0 ? 1u : 0u);
                bd += bstep;
              }
            }
          } else {
            // weight slices still arriving: the first tile of a resident set, or the ring (slot / phase counters, no divisions)
            auto tap_step = [&](int tap, uint32_t off16) {
              uint32_t slot;
              if (resident) {
                slot = (uint32_t)(c * taps + tap);
                mbar_wait_relaxed(&b_full[slot], 0u);
              } else {
                slot = bslot;
                TRACE_T0;
                mbar_wait_relaxed(&b_full[slot], bphase);
                TRACE_ADD(t_b);
              }
              tcgen05_fence_after();
              const uint64_t ad = adesc_s + (uint64_t)off16;
              const uint64_t bd = bdesc0 + (uint64_t)(slot * bstep);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (c | tap | k) != 0 ? 1u : 0u);
              if (!resident) {
                umma_commit(&b_empty[slot]);
                if (++bslot == (uint32_t)a.sb) { bslot = 0; bphase ^= 1u; }
              }
            };
            if (TAPS > 0) {
#pragma unroll
              for (int tap = 0; tap < NT; ++tap) tap_step(tap, toff[tap]);
            } else {
#pragma unroll 1
              for (int tap = 0; tap < taps; ++tap) tap_step(tap, tab[tap]);
            }
          }
          umma_commit(&a_empty[s]);
          ++ac;
          if (++astage == (uint32_t)a.sa) { astage = 0; aphase ^= 1u; }
        }
        umma_commit(&acc_full[buf]);
      }
#ifdef STM_CONV_TMA_TRACE
      if (blockIdx.x == 3) printf("conv_tma mma thread: %u tiles, total %lld clk, wait acc_empty %lld, a_full %lld, b_full %lld\n", tl, clock64() - t_all, t_acc, t_a, t_b);
#endif
    }
  } else {
    // =============================== EPILOGUE ===============================
    const int et = tid - 64;                       // 0..127
    const int q4 = warp & 3;                       // TMEM lane quarter this warp may read
    const int m = q4 * 32 + lane;                  // GEMM row = TMEM lane
    const bool relu = (a.flags & STM_DCN_RELU) != 0;
    const bool out_f32 = (a.flags & STM_DCN_OUT_F32) != 0;
    uint32_t tl = 0;
    int bias_n0 = -1;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tl) {
      const TileAt t = decode_tile(a, tile);
      const ConvTmaProb& q = a.prob[t.pi];
      const int n0 = t.nt * a.block_n;
      if (n0 != bias_n0) {                         // (uniform over the CTA's epilogue threads)
        named_barrier_sync(1, EPI_THREADS);        // everybody is done with the previous bias
        for (int i = et; i < a.block_n; i += EPI_THREADS) s_bias[i] = a.bias != nullptr ? __ldg(a.bias + n0 + i) : 0.f;
        named_barrier_sync(1, EPI_THREADS);
        bias_n0 = n0;
      }
      // row -> output pixel
      const int per_img = q.bh * q.bw;
      const int bi = m / per_img;
      const int rem = m - bi * per_img;
      const int r = rem / q.bw, cc = rem - r * q.bw;
      const int b = t.b0 + bi, ho = t.h0 + r, wo = t.w0 + cc;
      const bool ok = bi < q.bb && r < q.th && cc < q.tw && b < q.batch && ho < q.out_h && wo < q.out_w;
      const int64_t yoff = ok ? (b * q.y_sn + ho * q.y_sh + wo * q.y_sw + n0 * q.y_sc) : 0;
      const bool planar = q.y_sc != 1;
      const uint32_t buf = tl & 1u;
      if (a.exp_ & 4) mbar_wait_relaxed(&acc_full[buf], (tl >> 1) & 1u); else mbar_wait(&acc_full[buf], (tl >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * (uint32_t)a.n_mma;
      if (FUSE) {
        // column n * 3 + j holds P_j of output channel n for THIS row's pixel; out[m] = P_0[m] + P_1[m + 1] + P_2[m + 2]
        float* xq = s_xchg + ((tl & 1u) * 4 + q4) * 3 * a.block_n;            // what this quarter's rows 0 and 1 give the one below
        const float* xn = s_xchg + ((tl & 1u) * 4 + ((q4 + 1) & 3)) * 3 * a.block_n;
#pragma unroll 1
        for (int c0 = 0; c0 < a.block_n; c0 += 16) {
          uint32_t v[48];
          {
            uint32_t t0[16], t1[16], t2[16];
            tmem_ld16(taddr + (uint32_t)(c0 * 3), t0);
            tmem_ld16(taddr + (uint32_t)(c0 * 3 + 16), t1);
            tmem_ld16(taddr + (uint32_t)(c0 * 3 + 32), t2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = t0[i]; v[16 + i] = t1[i]; v[32 + i] = t2[i]; }
          }
          if (lane < 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (lane == 0) { xq[c0 + i] = __uint_as_float(v[3 * i + 1]); xq[a.block_n + c0 + i] = __uint_as_float(v[3 * i + 2]); }
              else xq[2 * a.block_n + c0 + i] = __uint_as_float(v[3 * i + 2]);
            }
          }
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[3 * i + 1]), 1);
            const float p2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[3 * i + 2]), 2);
            f[i] = __uint_as_float(v[3 * i]);
            if (lane < 31) f[i] += p1;
            if (lane < 30) f[i] += p2;
          }
          // rows m + 1 / m + 2 of the last two lanes live in the next quarter: wait for its boundary values
          // (both buffers of s_xchg alternate with the tile, so one barrier per 16 channels orders publish -> consume)
          named_barrier_sync(2, EPI_THREADS);
          if (lane >= 30 && q4 < 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (lane == 31) f[i] += xn[c0 + i] + xn[2 * a.block_n + c0 + i];
              else f[i] += xn[a.block_n + c0 + i];
            }
          }
          if (ok && !(a.exp_ & 2)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              f[i] += s_bias[c0 + i];
              if (relu) f[i] = fmaxf(f[i], 0.f);
            }
            if (planar) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int64_t o = yoff + (int64_t)(c0 + i) * q.y_sc;
                if (out_f32) reinterpret_cast<float*>(q.y)[o] = f[i];
                else reinterpret_cast<__nv_bfloat16*>(q.y)[o] = __float2bfloat16_rn(f[i]);
              }
            } else if (out_f32) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(q.y) + yoff + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(q.y) + yoff + c0);
              dst[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
              dst[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
            }
          }
        }
      } else {
#pragma unroll 1
      for (int c0 = 0; c0 < a.block_n; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (ok && !(a.exp_ & 2)) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = __uint_as_float(acc[i]) + s_bias[c0 + i];
            if (relu) f[i] = fmaxf(f[i], 0.f);
          }
          if (planar) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int64_t o = yoff + (int64_t)(c0 + i) * q.y_sc;
                if (out_f32) reinterpret_cast<float*>(q.y)[o] = f[i];
                else reinterpret_cast<__nv_bfloat16*>(q.y)[o] = __float2bfloat16_rn(f[i]);
              }
            } else if (out_f32) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(q.y) + yoff + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(q.y) + yoff + c0);
            dst[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            dst[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
          }
        }
      }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }

  // ---- teardown ----
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct TileCfg {
  int tw, th, bb, bw, bh;
  int64_t tiles;
  double eff;
};

// extra box columns / rows a tile needs beyond its own outputs
int halo_extent(int k, int pad, int stride) {
  if (stride == 1) return k - 1;
  const int t = k - 1 - pad;
  const int parity = t & 1;
  return ((t - parity) >> 1) + ((pad + 1) >> 1);
}

// TH x TW output pixels (x BB images when a whole map is smaller than a tile) with (TH - 1) * BW + TW <= 128 GEMM rows:
// the choice that leaves the fewest dead rows over the whole feature map
TileCfg choose_tile(int batch, int out_h, int out_w, int ex, int ey, int reserve = 0) {
  const int M = 128 - reserve;          // GEMM rows a tile's outputs may span (fused horizontal taps read `reserve` rows further)
  TileCfg best{};
  best.eff = -1.0;
  for (int tw = 1; tw <= out_w && tw <= 128; ++tw) {
    const int bw = tw + ex;
    if (tw > M) break;
    const int th = (M - tw) / bw + 1 < out_h ? (M - tw) / bw + 1 : out_h;
    const int bh = th + ey;
    int bb = 1;
    if (th == out_h && tw == out_w) bb = (M - ((th - 1) * bw + tw)) / (bh * bw) + 1;       // whole maps: several images per tile
    if (bw > 128 || bh > 128 || bb > 128) continue;
    // live rows per tile as if the batch were a multiple of bb: which kernel runs (this one or the gather loop, whose
    // accumulation order differs) must not depend on how a caller chunks its frames
    const double eff = (double)bb * out_h * out_w / ((double)((out_w + tw - 1) / tw) * ((out_h + th - 1) / th) * 128.0);
    // ties: the smaller box (less halo re-read)
    if (eff > best.eff + 1e-9 || (eff > best.eff - 1e-9 && bb * bh * bw < best.bb * best.bh * best.bw)) best = TileCfg{tw, th, bb, bw, bh, 0, eff};
  }
  if (best.eff > 0.0) {
    if (best.bb > batch) best.bb = batch > 0 ? batch : 1;
    best.tiles = (int64_t)((out_w + best.tw - 1) / best.tw) * ((out_h + best.th - 1) / best.th) * ((batch + best.bb - 1) / best.bb);
  }
  return best;
}

struct ConvTmaPlan {
  ConvTmaArgs args;
  int grid, smem_bytes, stride;
  double eff;
};

int pick_block_n_tma(int out_c) {
  if (out_c <= 256) return out_c;
  if (out_c % 256 == 0) return 256;
  if (out_c % 192 == 0) return 192;
  if (out_c % 128 == 0) return 128;
  return 0;
}

SmemAttrCache g_conv_tma_attr[6];

// horizontal taps fused into N: stride 1, 3x3, one N tile whose kw-fold still fits one UMMA (N <= 256) — the predictors
bool fuse_kw_ok(const StmDcnConv* c) {
  return c->stride_h == 1 && c->kernel_h == 3 && c->kernel_w == 3 && c->out_c * 3 <= 96 && (c->flags & STM_DCN_HINT_NO_FUSE) == 0;
}

}  // namespace

// Is this plain convolution one the TMA kernel runs (and runs well)?  A pure function of the arguments.
bool conv_tma_shape_supported(const StmDcnConv* c, const DcnParams& p, const char** why) {
  *why = "";
  if ((c->flags & STM_DCN_ZERO_OFFSET) == 0) { *why = "not a plain convolution"; return false; }
  if ((c->flags & STM_DCN_HINT_GATHER) != 0) { *why = "gather path requested"; return false; }
  if (c->dtype != STM_BF16 || c->groups != 1) { *why = "bf16, groups == 1 only"; return false; }
  if (c->dil_h != 1 || c->dil_w != 1) { *why = "dilation"; return false; }
  if (c->stride_h != c->stride_w || (c->stride_h != 1 && c->stride_h != 2)) { *why = "stride"; return false; }
  if (c->in_c % 64 != 0 || c->out_c % 16 != 0 || pick_block_n_tma(c->out_c) == 0) { *why = "channels"; return false; }
  if (c->kernel_h * c->kernel_w > 25 || c->pad_h >= c->kernel_h || c->pad_w >= c->kernel_w) { *why = "kernel / padding"; return false; }
  const int ex = halo_extent(c->kernel_w, c->pad_w, c->stride_w), ey = halo_extent(c->kernel_h, c->pad_h, c->stride_h);
  double rows = 0, live = 0;
  for (int i = 0; i < p.n_probs; ++i) {
    const DcnProblemDev& q = p.prob[i];
    const bool planar = (c->flags & STM_DCN_OUT_PLANAR) != 0;
    if (((uintptr_t)q.x & 15) || ((q.x_sn | q.x_sh | q.x_sw) & 7)) { *why = "alignment"; return false; }
    if (!planar && (((uintptr_t)q.y & 15) || ((q.y_sn | q.y_sh | q.y_sw) & 7))) { *why = "alignment"; return false; }
    if (q.x_sn < 0 || q.x_sh < 0 || q.x_sw < 0) { *why = "negative stride"; return false; }
    const TileCfg t = choose_tile(q.batch, q.out_h, q.out_w, ex, ey, fuse_kw_ok(c) ? c->kernel_w - 1 : 0);
    if (t.eff <= 0.0) { *why = "no tile"; return false; }
    const double px = (double)q.batch * q.out_h * q.out_w;
    rows += px / t.eff;
    live += px;
  }
  // dead GEMM rows cost tensor time the gather kernel does not spend: below ~55 % live rows it wins
  if (rows <= 0 || live / rows < 0.55) { *why = "tiles too sparse"; return false; }
  return true;
}

static int make_conv_tma_plan(const StmDcnConv* conv, const DcnParams& p, ConvTmaPlan* out) {
  ConvTmaPlan pl{};
  ConvTmaArgs& a = pl.args;
  const int S = p.sh;
  pl.stride = S;
  a.n_probs = p.n_probs;
  a.block_n = pick_block_n_tma(p.out_c);
  a.n_tiles = p.out_c / a.block_n;
  a.chunks = p.in_c / 64;
  a.kh = p.kh; a.kw = p.kw; a.ph = p.ph; a.pw = p.pw;
  a.flags = p.flags;
  a.exp_ = (conv->flags >> 20) & 0xff;
  a.bias = p.bias;
  const int ex = halo_extent(p.kw, p.pw, S), ey = halo_extent(p.kh, p.ph, S);
  a.fuse_kw = fuse_kw_ok(conv) ? p.kw : 0;
  a.n_mma = a.fuse_kw ? a.block_n * a.fuse_kw : a.block_n;
  int64_t sp_tiles = 0, live = 0;
  int phase_rows = 0;
  for (int i = 0; i < p.n_probs; ++i) {
    const DcnProblemDev& q = p.prob[i];
    const TileCfg t = choose_tile(q.batch, q.out_h, q.out_w, ex, ey, a.fuse_kw ? a.fuse_kw - 1 : 0);
    ConvTmaProb& d = a.prob[i];
    d.y = q.y; d.y_sn = q.y_sn; d.y_sh = q.y_sh; d.y_sw = q.y_sw;
    d.y_sc = (p.flags & STM_DCN_OUT_PLANAR) ? (int64_t)q.out_h * q.y_sh : 1;
    d.batch = q.batch; d.out_h = q.out_h; d.out_w = q.out_w;
    d.tw = t.tw; d.th = t.th; d.bb = t.bb; d.bw = t.bw; d.bh = t.bh;
    d.tiles_x = (q.out_w + t.tw - 1) / t.tw;
    d.tiles_y = (q.out_h + t.th - 1) / t.th;
    d.tiles_b = (q.batch + t.bb - 1) / t.bb;
    d.tile_begin = (int32_t)sp_tiles;
    d.box_bytes = t.bb * t.bh * t.bw * 128;
    sp_tiles += t.tiles;
    live += (int64_t)q.batch * q.out_h * q.out_w;
    const int rows = 128 + ey * t.bw + ex;                 // rows an M = 128 view can reach from the largest shift
    if (rows > phase_rows) phase_rows = rows;
  }
  if (sp_tiles * a.n_tiles >= (1ll << 31)) { set_error("tma conv: too many tiles"); return STM_ERR_UNSUPPORTED; }
  a.total_tiles = (int32_t)(sp_tiles * a.n_tiles);
  pl.eff = sp_tiles ? (double)live / ((double)sp_tiles * 128.0) : 0.0;
  a.a_phase_bytes = (phase_rows * 128 + 1023) & ~1023;
  a.a_stage_bytes = a.a_phase_bytes * (S == 2 ? 4 : 1);
  const int b_slot = a.n_mma * 128;
  const int fixed = (2 * MAX_SA + 2 * MAX_SB + 4) * 8 + 16 + a.block_n * 4 + STM_DCN_MAX_PROBLEMS * MAX_TAPS * 4 +
                    (a.fuse_kw ? 2 * 4 * 3 * a.block_n * 4 : 0) + 1024;
  const int views = a.fuse_kw ? p.kh : p.kh * p.kw;          // A views = weight slices per chunk
  const int items = a.chunks * views;
  a.sa = 2;
  int room = SMEM_LIMIT - fixed - a.sa * a.a_stage_bytes;
  int sb = room / b_slot;
  if (sb > MAX_SB) sb = MAX_SB;
  if (sb < 2) { set_error("tma conv: shared memory"); return STM_ERR_UNSUPPORTED; }
  a.resident = (a.n_tiles == 1 && items <= sb) ? 1 : 0;
  if (a.resident) {
    sb = items;
  } else {
    // a ring: one chunk's worth of taps in flight is plenty; spend what is left on a third / fourth input stage
    const int want = views + 3 < 12 ? 12 : views + 3;
    if (sb > want) sb = want;
  }
  a.sb = sb;
  room = SMEM_LIMIT - fixed - a.sa * a.a_stage_bytes - sb * b_slot;
  // a stage lives for only kh (fused) .. kh*kw tap views: the TMA round trip is hidden by the number of stages in flight
  while (a.sa < MAX_SA && room >= a.a_stage_bytes) { ++a.sa; room -= a.a_stage_bytes; }
  pl.smem_bytes = fixed + a.sa * a.a_stage_bytes + sb * b_slot;
  int cols = 32;
  while (cols < 2 * a.n_mma) cols <<= 1;
  a.tmem_cols = cols;
  const int sms = device_sm_count();
  pl.grid = a.total_tiles < sms ? a.total_tiles : sms;
  *out = pl;
  return STM_OK;
}

int conv_tma_variant(const StmDcnConv* conv, const DcnParams& p, char* buf, size_t len) {
  ConvTmaPlan pl;
  const int rc = make_conv_tma_plan(conv, p, &pl);
  if (rc != STM_OK) return rc;
  const ConvTmaArgs& a = pl.args;
  snprintf(buf, len, "tcgen05 tma-conv stride=%d n=%d fused_taps=%d tile=%dx%dx%d live=%.2f plain=1 a_stages=%d w_slots=%d resident=%d grid=%d tiles=%d",
           pl.stride, a.block_n, a.fuse_kw, a.prob[0].bb, a.prob[0].th, a.prob[0].tw, pl.eff, a.sa, a.sb, a.resident, pl.grid, a.total_tiles);
  return STM_OK;
}

int launch_conv_tma(const StmDcnConv* conv, const DcnParams& p, cudaStream_t stream) {
  ConvTmaPlan pl;
  const int rc = make_conv_tma_plan(conv, p, &pl);
  if (rc != STM_OK) return rc;
  if (pl.args.total_tiles == 0) return STM_OK;
  PFN_stm_encodeTiled enc = get_tensormap_encoder();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled unavailable"); return STM_ERR_CUDA; }
  ConvTmaMaps maps;
  const int S = pl.stride;
  for (int i = 0; i < p.n_probs; ++i) {
    const DcnProblemDev& q = p.prob[i];
    const ConvTmaProb& d = pl.args.prob[i];
    const cuuint64_t dims[4] = {(cuuint64_t)p.in_c, (cuuint64_t)q.in_w, (cuuint64_t)q.in_h, (cuuint64_t)q.batch};
    const cuuint64_t strides[3] = {(cuuint64_t)q.x_sw * 2, (cuuint64_t)q.x_sh * 2, (cuuint64_t)q.x_sn * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)(d.bw * S), (cuuint32_t)(d.bh * S), (cuuint32_t)d.bb};
    const cuuint32_t estr[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
    const CUresult r = enc(&maps.x[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(q.x), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tma conv: cuTensorMapEncodeTiled(x) failed (%d)", (int)r); return STM_ERR_UNSUPPORTED; }
  }
  for (int i = p.n_probs; i < STM_DCN_MAX_PROBLEMS; ++i) maps.x[i] = maps.x[0];
  if (pl.args.fuse_kw) {
    // OHWI weight [n][i][j][c] as a 4-D tensor (c, j, i, n): the box {64, kw, 1, N} of vertical tap i lands as rows n * kw + j
    const cuuint64_t dims[4] = {(cuuint64_t)p.in_c, (cuuint64_t)p.kw, (cuuint64_t)p.kh, (cuuint64_t)p.out_c};
    const cuuint64_t strides[3] = {(cuuint64_t)p.in_c * 2, (cuuint64_t)p.kw * p.in_c * 2, (cuuint64_t)p.kh * p.kw * p.in_c * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)p.kw, 1, (cuuint32_t)pl.args.block_n};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tma conv: cuTensorMapEncodeTiled(w, fused) failed (%d)", (int)r); return STM_ERR_UNSUPPORTED; }
  } else {
    const cuuint64_t ktot = (cuuint64_t)p.kh * p.kw * p.in_c;
    const cuuint64_t dims[2] = {ktot, (cuuint64_t)p.out_c};
    const cuuint64_t strides[1] = {ktot * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)pl.args.block_n};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p.w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tma conv: cuTensorMapEncodeTiled(w) failed (%d)", (int)r); return STM_ERR_UNSUPPORTED; }
  }
#define STM_CONV_GO(S_, T_, F_, SLOT_)                                                                             \
  do {                                                                                                             \
    const int rc2 = ensure_dynamic_smem(conv_tma_kernel<S_, T_, F_>, pl.smem_bytes, g_conv_tma_attr[SLOT_]);      \
    if (rc2 != STM_OK) return rc2;                                                                                 \
    conv_tma_kernel<S_, T_, F_><<<pl.grid, NUM_THREADS, pl.smem_bytes, stream>>>(pl.args, maps);                  \
  } while (0)
  const bool k33 = p.kh == 3 && p.kw == 3;
  if (pl.args.fuse_kw) { if (pl.args.exp_ & 16) STM_CONV_GO(1, 0, true, 5); else STM_CONV_GO(1, 3, true, 4); }
  else if (S == 1) { if (k33) STM_CONV_GO(1, 9, false, 0); else STM_CONV_GO(1, 0, false, 1); }
  else { if (k33) STM_CONV_GO(2, 9, false, 2); else STM_CONV_GO(2, 0, false, 3); }
#undef STM_CONV_GO
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace stm
