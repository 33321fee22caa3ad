// Deformable convolution forward as a bilinear-gather implicit GEMM on tcgen05 tensor cores.
//
//   Y[M, N] = A[M, K] * W[N, K]^T      M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin
//   A[m, (tap, c)] = mask[m, tap] * bilinear(x[b, :, :, c], p(m) + tap + offset[m, tap])
//
// A is never written to global memory.  Per CTA (256 or 128 output pixels x one N tile of <= 256):
//   * warps 0-15 PRODUCERS: once per (tap, deformable group) compute the four bilinear corner
//                weights/offsets of every row (DCN border rule, mask folded in) into shared memory;
//                then per 64-channel K block gather 4 x 16-byte corner vectors per (row, 8 channels)
//                (NHWC => contiguous), blend with fp32 accumulation (FHFMA.BF16: bf16 corner weights, no
//                unpack instructions), round once to bf16 and store into the
//                128B-swizzled K-major A tile; fence.proxy.async + mbarrier arrive.
//                After the main loop the same warps run the EPILOGUE: tcgen05.ld the fp32
//                accumulators, + bias, ReLU, bf16, 16-byte stores to NHWC y.
//   * warp 16    TMA: the matching [N x 64] slice of the packed OHWI weight -> swizzled B tile.
//   * warp 17    MMA: one thread issues tcgen05.mma (M=128, N<=256, K=16, bf16 -> fp32 in TMEM),
//                tcgen05.commit frees the stage; owns the TMEM allocation.
// Several feature maps that share a weight (FPN levels) are tiles of ONE launch.
//
// Replaces modulated_deformable_im2col + per-sample SGEMM of dcn_v2 / mmcv (reference
// backbone.py:45, Featurealign.py:72).
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace stm {

PFN_stm_encodeTiled get_tensormap_encoder() {
  static PFN_stm_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_stm_encodeTiled)p;
    else
      (void)cudaGetLastError();
  });
  return fn;
}

namespace {

using namespace tc;

constexpr int BLOCK_K = 64;             // bf16 elements = one 128-byte swizzle row
constexpr int TILE_M = 128;             // rows per accumulator (UMMA M)
constexpr int A_TILE_BYTES = TILE_M * 128;
constexpr int PRODUCER_WARPS = 16;
constexpr int PRODUCER_THREADS = PRODUCER_WARPS * 32;
constexpr int NUM_THREADS = PRODUCER_THREADS + 64;
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 227 * 1024;

struct TcArgs {
  DcnParams p;
  int32_t block_n;     // N tile (<= 256, multiple of 16)
  int32_t stages;
  int32_t tmem_cols;   // power of two >= M_TILES * block_n
  int32_t pad_;
};

template <int M_TILES>
struct SmemLayout {
  static constexpr int ROWS = TILE_M * M_TILES;
  int stage_bytes, meta_w, meta_o, row_x, row_y, row_pos, bars, total;
  __host__ __device__ SmemLayout(int block_n, int stages) {
    stage_bytes = M_TILES * A_TILE_BYTES + block_n * 128;
    int off = stages * stage_bytes;
    meta_w = off; off += 2 * ROWS * 16;
    meta_o = off; off += 2 * ROWS * 16;
    row_x = off;  off += ROWS * 8;   // (kept 8 B/row: uint32 batch offset in elements + pad)
    row_y = off;  off += ROWS * 8;
    row_pos = off; off += ROWS * 16;
    bars = off;   off += (2 * MAX_STAGES + 2) * 8;
    total = off + 1024;   // slack for manual 1024-byte alignment of the base
  }
};

template <int M_TILES, int DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dcn_tc_kernel(const __grid_constant__ TcArgs a, const __grid_constant__ CUtensorMap tmap_w) {
  constexpr int ROWS = TILE_M * M_TILES;
  constexpr int ROWS_PER_PASS = PRODUCER_WARPS * 4;     // 64 rows per sweep of the producer warps
  constexpr int PASSES = ROWS / ROWS_PER_PASS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout<M_TILES> L(a.block_n, a.stages);
  uint4* meta_w = reinterpret_cast<uint4*>(smem + L.meta_w);    // {w0|w1, w2|w3 as bf16 pairs, nonzero-corner bits, -}
  int4* meta_o = reinterpret_cast<int4*>(smem + L.meta_o);
  uint32_t* row_x = reinterpret_cast<uint32_t*>(smem + L.row_x);     // element offset of x[b, 0, 0, 0] from the problem base
  int64_t* row_y = reinterpret_cast<int64_t*>(smem + L.row_y);
  int4* row_pos = reinterpret_cast<int4*>(smem + L.row_pos);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* accum_bar = empty_bar + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const DcnParams& p = a.p;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int block_n = a.block_n, stages = a.stages;
  const int n0 = blockIdx.y * block_n;

  // ---- which feature map does this M block belong to? ----
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < p.n_probs; ++i)
    if ((int)blockIdx.x >= p.prob[i].tile_begin) pi = i;
  const DcnProblemDev& pr = p.prob[pi];
  const int m0 = ((int)blockIdx.x - pr.tile_begin) * ROWS;

  // ---- one-time setup ----
  if (tid < ROWS) {
    const int m = m0 + tid;
    int b = 0, ho = 0, wo = 0, valid = 0;
    int64_t yb = -1;
    if (m < pr.m_total) {
      const int hw = pr.out_h * pr.out_w;
      b = m / hw;
      const int r = m - b * hw;
      ho = r / pr.out_w;
      wo = r - ho * pr.out_w;
      yb = b * pr.y_sn + ho * pr.y_sh + wo * pr.y_sw;
      valid = 1;
    }
    row_x[tid] = (uint32_t)(b * pr.x_sn);
    row_y[tid] = yb;
    row_pos[tid] = make_int4(b, ho, wo, valid);
  }
  if (warp == PRODUCER_WARPS && lane == 0) {
    prefetch_tensormap(&tmap_w);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], PRODUCER_WARPS + 1);   // producer warps + the TMA thread's expect_tx arrive
      mbar_init(&empty_bar[s], 1);                   // tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == PRODUCER_WARPS + 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int K = p.kh * p.kw;
  const int cpd = p.in_c / p.dg;            // channels per deformable group (multiple of 64)
  const int chunks = cpd / BLOCK_K;

  if (warp < PRODUCER_WARPS) {
    // =============================== PRODUCERS ===============================
    const int v = lane & 7;                 // 16-byte (8-channel) slot inside the 64-channel K block
    const int rsub = lane >> 3;             // 4 rows per warp instruction
    const int row0 = warp * 4 + rsub;       // this thread's row in pass 0; pass j adds j * ROWS_PER_PASS
    // 128B swizzle: chunk v of row r goes to chunk v ^ (r & 7); r & 7 is the same in every pass
    const int swz = (v ^ (row0 & 7)) << 4;
    const bool has_off = pr.offset != nullptr, has_mask = pr.mask != nullptr;
    const bool off_bf16 = (p.flags & 0x100) != 0;      // internal flag: offsets/masks stored as bf16
    const int4 pos = (tid < ROWS) ? row_pos[tid] : make_int4(0, 0, 0, 0);
    const __nv_bfloat16* __restrict__ xbase = reinterpret_cast<const __nv_bfloat16*>(pr.x);
    const int n_iter = K * p.dg;

    // raw (dy, dx, mask) of this thread's row for iteration `it`; issued one iteration ahead so the
    // global-load latency hides behind the gather instead of stalling everybody at the named barrier
    auto load_raw = [&](int it_, float& oy, float& ox, float& mk) {
      oy = 0.f; ox = 0.f; mk = 1.f;
      if (tid >= ROWS || !pos.w || it_ >= n_iter) return;
      const int tap_ = it_ / p.dg, g_ = it_ - tap_ * p.dg;
      if (has_off) {
        const int64_t o = pos.x * pr.off_sn + (int64_t)(g_ * 2 * K + 2 * tap_) * pr.off_sc + pos.y * pr.off_sh + pos.z * pr.off_sw;
        if (off_bf16) {
          oy = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pr.offset)[o]);
          ox = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pr.offset)[o + pr.off_sc]);
        } else {
          oy = reinterpret_cast<const float*>(pr.offset)[o];
          ox = reinterpret_cast<const float*>(pr.offset)[o + pr.off_sc];
        }
      }
      if (has_mask) {
        const int64_t o = pos.x * pr.mask_sn + (int64_t)(g_ * K + tap_) * pr.mask_sc + pos.y * pr.mask_sh + pos.z * pr.mask_sw;
        mk = off_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pr.mask)[o])
                      : reinterpret_cast<const float*>(pr.mask)[o];
      }
    };

    // One gather task = (row, 8 channels) of one K block: 4 corner loads, fp32 blend, one 16-byte store.
    // Two register sets (A, B) rotate so the loads of task t+1 are in flight while task t is blended.
    struct GTask {
      uint32_t w01, w23;   // bf16 corner weights (w0 | w1 << 16, w2 | w3 << 16)
      uint32_t dst;        // shared-memory address of this task's 16-byte slot in the A tile
      uint4 c[4];
    };
    auto issue = [&](GTask& t, int buf, int row, int chan, uint32_t a_stage) {
      const uint4 mw = meta_w[buf * ROWS + row];
      const int4 o4 = meta_o[buf * ROWS + row];
      const uint32_t eb = row_x[row] + (uint32_t)chan;     // 32-bit element offsets from the problem base (host checks < 2^31)
      t.w01 = mw.x; t.w23 = mw.y;
      t.dst = a_stage + row * 128 + swz;                 // row r of the (stacked) A tiles lives at r * 128
      const uint4* s0 = reinterpret_cast<const uint4*>(xbase + (eb + (uint32_t)o4.x));
      const uint4* s1 = reinterpret_cast<const uint4*>(xbase + (eb + (uint32_t)o4.y));
      const uint4* s2 = reinterpret_cast<const uint4*>(xbase + (eb + (uint32_t)o4.z));
      const uint4* s3 = reinterpret_cast<const uint4*>(xbase + (eb + (uint32_t)o4.w));
      // every lane's four corners carry weight (the common, interior case): plain loads.  Otherwise
      // zero-weight corners are NOT read, so data outside the sample can never leak in (0 * Inf).
      if (DBG == 1) {
        t.c[0] = t.c[1] = t.c[2] = t.c[3] = make_uint4((uint32_t)(uintptr_t)s0, (uint32_t)(uintptr_t)s1, (uint32_t)(uintptr_t)s2, (uint32_t)(uintptr_t)s3);
      } else if (__all_sync(0xffffffffu, mw.z == 15u)) {
        t.c[0] = __ldg(s0); t.c[1] = __ldg(s1); t.c[2] = __ldg(s2); t.c[3] = __ldg(s3);
      } else {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        t.c[0] = (mw.z & 1u) ? __ldg(s0) : z;
        t.c[1] = (mw.z & 2u) ? __ldg(s1) : z;
        t.c[2] = (mw.z & 4u) ? __ldg(s2) : z;
        t.c[3] = (mw.z & 8u) ? __ldg(s3) : z;
      }
    };
    auto finish = [&](const GTask& t) {
      const uint32_t* q0 = reinterpret_cast<const uint32_t*>(&t.c[0]);
      const uint32_t* q1 = reinterpret_cast<const uint32_t*>(&t.c[1]);
      const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&t.c[2]);
      const uint32_t* q3 = reinterpret_cast<const uint32_t*>(&t.c[3]);
      uint16_t w0, w1, w2, w3;
      split16(t.w01, w0, w1);
      split16(t.w23, w2, w3);
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint16_t a0, a1, b0, b1, c0, c1, d0, d1;
        split16(q0[i], a0, a1);
        split16(q1[i], b0, b1);
        split16(q2[i], c0, c1);
        split16(q3[i], d0, d1);
        float lo = fma_bf16(a0, w0, 0.f);
        float hi = fma_bf16(a1, w0, 0.f);
        lo = fma_bf16(b0, w1, lo);
        hi = fma_bf16(b1, w1, hi);
        lo = fma_bf16(c0, w2, lo);
        hi = fma_bf16(c1, w2, hi);
        lo = fma_bf16(d0, w3, lo);
        hi = fma_bf16(d1, w3, hi);
        o[i] = pack_bf16(lo, hi);
      }
      if (DBG == 3) { o[0] = q0[0]; o[1] = q0[1]; o[2] = q0[2]; o[3] = q0[3]; }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t.dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    };
    auto publish = [&](int stage) {        // this warp's part of the A tile of `stage` is complete
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    };

    GTask A, B;
    int stage = 0, prev_stage = 0;
    uint32_t phase = 0;
    bool pending = false;                   // B holds the last task of the previous K block
    float r_oy, r_ox, r_mk;
    load_raw(0, r_oy, r_ox, r_mk);
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
      const int tap = it / p.dg, g = it - tap * p.dg;
      const int ti = tap / p.kw, tj = tap - ti * p.kw;
      const int buf = it & 1;
      if (tid < ROWS) {
        Sample4 sm;
#pragma unroll
        for (int i = 0; i < 4; ++i) { sm.w[i] = 0.f; sm.o[i] = 0; }
        if (pos.w) {
          const float mk = (has_mask && (p.flags & STM_DCN_MASK_SIGMOID)) ? sigmoidf_(r_mk) : r_mk;
          const float h = (float)(pos.y * p.sh - p.ph + ti * p.dh) + r_oy;
          const float w = (float)(pos.z * p.sw - p.pw + tj * p.dw) + r_ox;
          sm = make_sample(h, w, pr.in_h, pr.in_w, pr.x_sh, pr.x_sw, mk);
        }
        {
          const uint32_t w01 = pack_bf16(sm.w[0], sm.w[1]), w23 = pack_bf16(sm.w[2], sm.w[3]);
          const uint32_t nz = ((w01 & 0x7fffu) ? 1u : 0u) | ((w01 & 0x7fff0000u) ? 2u : 0u) | ((w23 & 0x7fffu) ? 4u : 0u) |
                              ((w23 & 0x7fff0000u) ? 8u : 0u);
          meta_w[buf * ROWS + tid] = make_uint4(w01, w23, nz, 0u);
        }
        meta_o[buf * ROWS + tid] = make_int4(sm.o[0], sm.o[1], sm.o[2], sm.o[3]);
      }
      load_raw(it + 1, r_oy, r_ox, r_mk);
      // meta[buf] was last read two iterations ago; every thread has passed the previous barrier since
      named_barrier_sync(1, PRODUCER_THREADS);
#pragma unroll 1
      for (int cc = 0; cc < chunks; ++cc) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const uint32_t a_stage = smem_u32(smem + stage * L.stage_bytes);
        const int chan = g * cpd + cc * BLOCK_K + v * 8;
#pragma unroll
        for (int j = 0; j < PASSES; j += 2) {
          issue(A, buf, row0 + j * ROWS_PER_PASS, chan, a_stage);
          if (j == 0) {
            if (pending) { finish(B); publish(prev_stage); }
          } else {
            finish(B);
          }
          issue(B, buf, row0 + (j + 1) * ROWS_PER_PASS, chan, a_stage);
          finish(A);
        }
        pending = true;
        prev_stage = stage;
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (pending) { finish(B); publish(prev_stage); }
    // =============================== EPILOGUE ===============================
    mbar_wait(accum_bar, 0);
    tcgen05_fence_after();
    // warp w may only touch TMEM lanes [32 (w % 4), +32).  The four warp groups (w / 4) split the
    // accumulators: M_TILES == 2 -> (tile, column half); M_TILES == 1 -> column quarter.
    const int q = warp & 3;
    const int grp = warp >> 2;
    const int mt = (M_TILES == 2) ? (grp & 1) : 0;
    const int part = (M_TILES == 2) ? (grp >> 1) : grp;
    constexpr int PARTS = (M_TILES == 2) ? 2 : 4;
    const int nchunk = block_n / 16;
    const int c_begin = (part * nchunk / PARTS) * 16;
    const int c_end = ((part + 1) * nchunk / PARTS) * 16;
    const int row = mt * TILE_M + q * 32 + lane;
    const int64_t yoff = row_y[row];
    __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(pr.y) + (yoff >= 0 ? yoff : 0) + n0;
    const bool relu = (p.flags & STM_DCN_RELU) != 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * block_n);
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      uint32_t acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (yoff >= 0) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float lo = __uint_as_float(acc[2 * i]), hi = __uint_as_float(acc[2 * i + 1]);
          if (p.bias != nullptr) {
            lo += __ldg(p.bias + n0 + c0 + 2 * i);
            hi += __ldg(p.bias + n0 + c0 + 2 * i + 1);
          }
          if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
          o[i] = pack_bf16(lo, hi);
        }
        uint4* dst = reinterpret_cast<uint4*>(yrow + c0);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
  } else if (warp == PRODUCER_WARPS) {
    // =============================== TMA (weights) ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int tap = 0; tap < K; ++tap)
#pragma unroll 1
        for (int g = 0; g < p.dg; ++g)
#pragma unroll 1
          for (int cc = 0; cc < chunks; ++cc) {
            mbar_wait_relaxed(&empty_bar[s], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(block_n * 128));
            tma_load_2d(smem + s * L.stage_bytes + M_TILES * A_TILE_BYTES, &tmap_w, &full_bar[s],
                        tap * p.in_c + g * cpd + cc * BLOCK_K, n0);
            if (++s == stages) { s = 0; phase ^= 1u; }
          }
    }
  } else {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TILE_M, (uint32_t)block_n);
      const int num_kb = K * p.dg * chunks;
      int s = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_relaxed(&full_bar[s], phase);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L.stage_bytes);
        const uint64_t bdesc = umma_desc_sw128(a_addr + M_TILES * A_TILE_BYTES);
#pragma unroll
        for (int mt = 0; mt < M_TILES; ++mt) {
          const uint64_t adesc = umma_desc_sw128(a_addr + mt * A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            if (DBG != 2) umma_bf16(tmem_base + (uint32_t)(mt * block_n), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);          // stage reusable once these MMAs have read it
        if (++s == stages) { s = 0; phase ^= 1u; }
      }
      umma_commit(accum_bar);                // accumulators complete
    }
    __syncwarp();
  }

  // ---- teardown ----
  tcgen05_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARPS + 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int pick_block_n(int out_c) {
  if (out_c <= 256) return out_c;
  if (out_c % 256 == 0) return 256;
  if (out_c % 192 == 0) return 192;
  if (out_c % 128 == 0) return 128;
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = 148;
    }
  }
  return n;
}

template <int M_TILES, int DBG = 0>
int launch_t(const TcArgs& args, const CUtensorMap& tmap, dim3 grid, int smem_bytes, cudaStream_t stream) {
  static int configured = 0;
  if (configured < smem_bytes) {
    STM_CUDA_OK(cudaFuncSetAttribute(dcn_tc_kernel<M_TILES, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = smem_bytes;
  }
  dcn_tc_kernel<M_TILES, DBG><<<grid, NUM_THREADS, smem_bytes, stream>>>(args, tmap);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace

bool dcn_tc_supported(const StmDcnConv* c, const StmDcnProblem* pr, int n, const char** why) {
  *why = "";
  if (c->dtype != STM_BF16) { *why = "dtype is not bf16"; return false; }
  if (c->groups != 1) { *why = "groups != 1"; return false; }
  if (c->in_c % 64 != 0 || (c->in_c / c->deform_groups) % 64 != 0) { *why = "channels per deformable group not a multiple of 64"; return false; }
  if (c->out_c % 16 != 0 || pick_block_n(c->out_c) == 0) { *why = "out_c not tileable (multiple of 16, <= 256 or a multiple of 128)"; return false; }
  if (c->kernel_h * c->kernel_w > 64) { *why = "kernel too large"; return false; }
  for (int i = 0; i < n; ++i) {
    const StmDcnProblem& q = pr[i];
    if (q.batch == 0) continue;
    if (((uintptr_t)q.x & 15) || ((uintptr_t)q.y & 15)) { *why = "x / y not 16-byte aligned"; return false; }
    if ((int64_t)q.batch * q.x_stride_n >= (1ll << 31)) { *why = "x spans more than 2^31 elements"; return false; }
    if ((q.x_stride_n | q.x_stride_h | q.x_stride_w | q.y_stride_n | q.y_stride_h | q.y_stride_w) & 7) {
      *why = "x / y strides not multiples of 8 elements";
      return false;
    }
  }
  if (get_tensormap_encoder() == nullptr) { *why = "cuTensorMapEncodeTiled unavailable"; return false; }
  return true;
}

size_t dcn_tc_workspace(const StmDcnConv*, const StmDcnProblem*, int) { return 0; }

int launch_dcn_tc(const StmDcnConv* conv, const DcnParams& p_in, void*, size_t, cudaStream_t stream) {
  TcArgs args;
  args.p = p_in;
  DcnParams& p = args.p;
  if (conv->offset_dtype == STM_BF16) p.flags |= 0x100;
  const int block_n = pick_block_n(p.out_c);
  int64_t rows = 0;
  for (int i = 0; i < p.n_probs; ++i) rows += p.prob[i].m_total;
  const int n_tiles = p.out_c / block_n;
  // two accumulators (256 rows) per CTA halve the weight traffic per row (B200: 0.205 -> 0.150 ms on the 24x40
  // backbone layers); fall back to 128-row CTAs only when 256-row CTAs could not even half-fill the GPU
  const int64_t ctas256 = ((rows + 255) / 256) * n_tiles;
  int m_tiles = (ctas256 >= sm_count() / 2 && 2 * block_n <= 512) ? 2 : 1;
  if (const char* e = getenv("STM_DCN_MTILES")) {           // tuning knob (profiling runs)
    const int v = atoi(e);
    if ((v == 1 || v == 2) && v * block_n <= 512) m_tiles = v;
  }
  const int rows_per_cta = TILE_M * m_tiles;
  int blocks = 0;
  for (int i = 0; i < p.n_probs; ++i) {
    p.prob[i].tile_begin = blocks;
    blocks += (p.prob[i].m_total + rows_per_cta - 1) / rows_per_cta;
  }
  p.total_m_tiles = blocks;
  if (blocks == 0) return STM_OK;
  args.block_n = block_n;
  int cols = 32;
  while (cols < m_tiles * block_n) cols <<= 1;
  args.tmem_cols = cols;
  args.pad_ = 0;
  // pipeline depth: as many stages as fit while leaving L1 some room for the gather's corner reuse
  int stages = MAX_STAGES, smem_bytes = 0;
  int budget = m_tiles == 2 ? 164 * 1024 : 132 * 1024;
  if (const char* e = getenv("STM_DCN_SMEM_KB")) {          // tuning knob (profiling runs)
    const int v = atoi(e);
    if (v >= 64 && v <= 227) budget = v * 1024;
  }
  for (; stages >= 2; --stages) {
    smem_bytes = m_tiles == 2 ? SmemLayout<2>(block_n, stages).total : SmemLayout<1>(block_n, stages).total;
    if (smem_bytes <= budget) break;
  }
  if (stages < 2) {
    stages = 2;
    smem_bytes = m_tiles == 2 ? SmemLayout<2>(block_n, 2).total : SmemLayout<1>(block_n, 2).total;
  }
  if (smem_bytes > SMEM_LIMIT) { set_error("tcgen05 DCN: shared memory %d B over the limit", smem_bytes); return STM_ERR_UNSUPPORTED; }
  args.stages = stages;

  CUtensorMap tmap;
  const cuuint64_t ktot = (cuuint64_t)p.kh * p.kw * p.in_c;
  const cuuint64_t dims[2] = {ktot, (cuuint64_t)p.out_c};
  const cuuint64_t strides[1] = {ktot * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  const cuuint32_t estr[2] = {1, 1};
  PFN_stm_encodeTiled enc = get_tensormap_encoder();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled unavailable"); return STM_ERR_CUDA; }
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p.w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return STM_ERR_CUDA; }

  const dim3 grid((unsigned)blocks, (unsigned)n_tiles);
#ifdef STM_DCN_EXPERIMENTS
  if (const char* e = getenv("STM_DCN_DBG")) {          // bring-up experiments only (wrong results by design)
    const int v = atoi(e);
    if (m_tiles == 2 && v == 1) return launch_t<2, 1>(args, tmap, grid, smem_bytes, stream);
    if (m_tiles == 2 && v == 2) return launch_t<2, 2>(args, tmap, grid, smem_bytes, stream);
    if (m_tiles == 2 && v == 3) return launch_t<2, 3>(args, tmap, grid, smem_bytes, stream);
  }
#endif
  if (m_tiles == 2) return launch_t<2>(args, tmap, grid, smem_bytes, stream);
  return launch_t<1>(args, tmap, grid, smem_bytes, stream);
}

}  // namespace stm
