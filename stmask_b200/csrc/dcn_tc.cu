// Deformable convolution forward as a bilinear-gather implicit GEMM on tcgen05 tensor cores.
//
//   Y[M, N] = A[M, K] * W[N, K]^T      M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin
//   A[m, (tap, c)] = mask[m, tap] * bilinear(x[b, :, :, c], p(m) + tap + offset[m, tap])
//
// A is never written to global memory.  Per CTA (256 or 128 output pixels x one N tile of <= 256):
//   * producer warps (16, or 8 with a deeper look-ahead): one tap ahead of the gather, all threads compute
//                the four corner POINTERS (64-bit; corners without weight point at a zero page, so the
//                gather needs no predicates) and bf16 corner weights (DCN border rule, mask folded in) of
//                every row into triple-buffered shared memory.  Per 64-channel K block every thread runs a
//                register-pipelined stream of gather tasks: 4 x 16-byte corner loads (NHWC => contiguous),
//                fp32 blend (FHFMA.BF16: bf16 data and weights feed the FMA directly), one rounding to
//                bf16, 16-byte store into the 128B-swizzled K-major A tile; one mbarrier arrive per warp.
//                Afterwards the same warps are the EPILOGUE: tcgen05.ld the fp32 accumulators, + bias,
//                ReLU, bf16, 16-byte stores to NHWC y.
//   * TMA warp:  the matching [N x 64] slice of the packed OHWI weight -> swizzled B tile.
//   * MMA warp:  one thread issues fence.proxy.async + tcgen05.mma (M=128, N<=256, K=16, bf16 -> fp32 in
//                TMEM), tcgen05.commit frees the stage; owns the TMEM allocation.
// Several feature maps that share a weight (FPN levels) are tiles of ONE launch.
//
// Replaces modulated_deformable_im2col + per-sample SGEMM of dcn_v2 / mmcv (reference
// backbone.py:45, Featurealign.py:72).
#include <cstdio>
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
constexpr int MAX_STAGES = 6;
constexpr int META_BUFS = 3;            // sample metadata is computed one tap ahead of the gather that reads it
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int ZERO_PAGE_BYTES = 4096;   // >= 2 * in_c (dcn_tc_supported caps in_c at 2048)
constexpr int FLAG_OFFSETS_BF16 = 1 << 16;   // internal (set by the launcher): offsets / masks are stored as bf16

// Corners that carry no weight (outside the map, outside the sample's support, rows past the end of the
// problem) are pointed here instead of being predicated off: the gather stays branch-free and 0 * 0 = 0,
// so nothing outside a sample's support can leak into it.
__device__ __align__(128) unsigned char g_zero_page[ZERO_PAGE_BYTES + 128];

struct TcArgs {
  DcnParams p;
  int32_t block_n;     // N tile (<= 256, multiple of 16)
  int32_t stages;
  int32_t tmem_cols;   // power of two >= M_TILES * block_n
  int32_t pad_;
  int32_t chunk_outer; // K-block order (chunk, tap) with every tap's sample records resident, instead of (tap, chunk)
  int32_t pad2_;
};

template <int M_TILES>
struct SmemLayout {
  static constexpr int ROWS = TILE_M * M_TILES;
  int stage_bytes, meta_p, meta_w, bars, fcb_w, total;
  // pair: each CTA of a cta_group::2 pair holds only its half of the N tile's weight rows;
  // fcb_floats: FCB(ada) conv_offset weights [dg * 2K][4] kept in shared memory
  // meta_bufs: sample-record buffers (3 when the records are computed one tap ahead; all kh*kw taps when resident)
  __host__ __device__ SmemLayout(int block_n, int stages, bool pair, int fcb_floats = 0, int meta_bufs = META_BUFS) {
    stage_bytes = M_TILES * A_TILE_BYTES + (pair ? block_n / 2 : block_n) * 128;
    int off = stages * stage_bytes;
    meta_p = off; off += meta_bufs * ROWS * 16;    // per row: top-left corner pointer + 2 flag bits, four bf16 corner weights
    meta_w = off;
    bars = off;   off += (2 * MAX_STAGES + 2) * 8;
    fcb_w = off;  off += (fcb_floats * 4 + 15) & ~15;
    total = off + 1024;   // slack for manual 1024-byte alignment of the base
  }
};

// GEMM row r of one image -> output pixel.  Raster order, or 8x8 PATCH order (maps whose sides are multiples of 8): a
// CTA's 128 rows are then two neighbouring 8x8 patches instead of 1.6 image rows, its taps' footprint in the input is a
// ~17 x 25 pixel window instead of ~10 x 88, and what the gather misses in L1 and fetches from L2 shrinks with it — the
// step runs at the 1 kW power cap, where L2 traffic is clock.
__device__ __forceinline__ void row_to_pixel(const DcnProblemDev& pr, int r, int& ho, int& wo) {
  if (pr.patch) {
    const int p = r >> 6, w = r & 63;
    const int pw = pr.out_w >> 3;
    const int py = p / pw, px = p - py * pw;
    ho = py * 8 + (w >> 3);
    wo = px * 8 + (w & 7);
  } else {
    ho = r / pr.out_w;
    wo = r - ho * pr.out_w;
  }
}

__device__ __forceinline__ uint4 ldg_nc_v4(uint64_t addr) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(addr));
  return r;
}
__device__ __forceinline__ uint64_t add_wide(uint32_t lo, uint32_t hi, uint32_t off) {
  uint64_t r;
  const uint64_t base = ((uint64_t)hi << 32) | lo;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(base));
  return r;
}

// PW producer warps (8 or 16) + 1 TMA warp + 1 MMA warp.  The generic->async proxy fence for the A tile is
// executed by the MMA thread after it has acquired the stage (instead of by every producer thread before its
// release): a producer-side fence.proxy.async compiles to MEMBAR.ALL.CTA, which also waits for the gather
// loads that are already in flight for the NEXT K block and drains the pipeline.
//
// PAIR: two CTAs of a cluster (same TPC) work as ONE tcgen05 cta_group::2 unit.  Each CTA gathers the A tile of
// its own ROWS output pixels and loads only HALF of the N-tile's weight rows; the leader's MMA (M = 256: 128
// TMEM lanes in each CTA) reads both halves.  Per CTA that halves the weight bytes fetched from L2 and the
// shared-memory bytes the tensor core reads for the B operand (B is 2/3 of the operand traffic at N = 256) —
// the shared-memory/L1 data pipe is what bounds this kernel (gather loads, A-tile stores and UMMA operand
// reads all go through it).  Synchronisation: every CTA's producer warps arrive on their OWN full barrier.  In
// the leader that barrier also collects the weight TMA of BOTH CTAs (the peer's TMA completes its bytes on the
// leader's barrier) and one arrival per K block from the peer's RELAY thread (the peer's otherwise idle MMA
// warp): it waits for the peer's A tile, runs the generic->async proxy fence ON THE PEER SM (whose tensor core
// reads that tile) and forwards one remote arrive.  Data never crosses CTAs through the generic proxy, only
// that signal does, so all barrier operations keep CTA scope (cluster-scope release / acquire compile to
// MEMBAR.ALL.GPU / CCTL.IVALL per K block — measured 40 % slower than no pairing at all).  The leader's
// tcgen05.commit multicasts to the empty barriers of both CTAs.
//
// PLAIN: zero offsets (STM_DCN_ZERO_OFFSET): the sample is the pixel itself, so a gather task is ONE 16-byte
// load and a store (no blend) and the metadata is one pointer per row.  This is the regular convolution that
// predicts a DCN's offsets / mask logits (backbone.py:24-26) and the TemporalNet convs
// (track_to_segment_head.py:14-16) on the same tcgen05 main loop.
//
// MODE 2 (FCB, box-guided offsets): `offset` is not the per-tap offset tensor but the four regressed box deltas
// (t_x, t_y, t_w, t_h) per pixel; the producers derive every tap's (dy, dx) themselves while they compute the
// sampling metadata — FCB(ada): the 1x1 `conv_offset` (8 FMAs per tap, weights in shared memory), FCB(ali): the closed
// form of Featurealign.py:46-69.  The 30-channel offset tensors and the kernels that wrote them disappear.
template <int M_TILES, int PW, int D, bool PAIR, int MODE>
__global__ void __launch_bounds__(PW * 32 + 64, (M_TILES == 1 && PW == 8) ? 2 : 1)
dcn_tc_kernel(const __grid_constant__ TcArgs a, const __grid_constant__ CUtensorMap tmap_w) {
  constexpr bool PLAIN = MODE == 1;
  constexpr bool FCB = MODE == 2;
  constexpr int ROWS = TILE_M * M_TILES;
  constexpr int PT = PW * 32;                           // producer threads
  constexpr int ROWS_PER_PASS = PW * 4;                 // rows covered by one sweep of the producer warps
  constexpr int TPK = ROWS / ROWS_PER_PASS;             // gather tasks per thread per K block
  constexpr int CPT = 4 * ROWS / PT;                    // corners per thread when computing sample metadata
  constexpr int NC = PLAIN ? 1 : 4;                     // corner loads per gather task
  static_assert(CPT == 1 || CPT == 2 || CPT == 4, "metadata split");
  static_assert(D >= 1 && D <= TPK && TPK % D == 0, "gather look-ahead must divide the tasks per K block");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#ifdef STM_DCN_TRACE
  long long tr_[8];
  tr_[0] = clock64();
#define STM_TR(i) tr_[i] = clock64()
#else
#define STM_TR(i)
#endif
  const bool CO = !PLAIN && a.chunk_outer != 0;      // (chunk, tap) K-block order, all taps' records resident (dg == 1)
  const SmemLayout<M_TILES> L(a.block_n, a.stages, PAIR, FCB ? a.p.dg * 2 * a.p.kh * a.p.kw * 4 : 0, CO ? a.p.kh * a.p.kw : META_BUFS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* accum_bar = empty_bar + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const DcnParams& p = a.p;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int block_n = a.block_n, stages = a.stages;
  const int n0 = blockIdx.y * block_n;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;

  // ---- which feature map does this M block belong to?  (a pair's odd filler CTA past the last tile owns no rows) ----
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < p.n_probs; ++i)
    if ((int)blockIdx.x >= p.prob[i].tile_begin) pi = i;
  const DcnProblemDev& pr = p.prob[pi];
  const int m0 = ((int)blockIdx.x - pr.tile_begin) * ROWS;

  // ---- one-time setup ----
  if (warp == PW && lane == 0) {
    prefetch_tensormap(&tmap_w);
    for (int s = 0; s < stages; ++s) {
      // producer warps + the TMA thread's expect_tx arrive (+ the peer's relay thread in a pair's leader);
      // a pair's non-leader: its producer warps only (its relay thread is the waiter)
      mbar_init(&full_bar[s], !PAIR ? PW + 1 : (leader ? PW + 2 : PW));
      mbar_init(&empty_bar[s], 1);                   // tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (FCB && (p.flags & STM_DCN_FCB_ADA) && warp < PW) {
    float* fw = reinterpret_cast<float*>(smem + L.fcb_w);
    for (int i = tid; i < p.dg * 2 * p.kh * p.kw * 4; i += PW * 32) fw[i] = __ldg(p.fcb_w + i);
  }
  if (warp == PW + 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, (uint32_t)a.tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols); tmem_relinquish(); }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the leader's full barriers as seen from this CTA
  const uint32_t full_leader = PAIR ? map_to_cta(smem_u32(full_bar), 0u) : smem_u32(full_bar);

  const int K = p.kh * p.kw;
  const int cpd = p.in_c / p.dg;            // channels per deformable group (multiple of 64)
  const int chunks = cpd / BLOCK_K;
  const int n_iter = K * p.dg;              // (tap, deformable group) pairs, tap-major
  const int num_kb = n_iter * chunks;

  if (warp < PW) {
    // =============================== PRODUCERS ===============================
    const int v = lane & 7;                 // 16-byte (8-channel) slot inside the 64-channel K block
    const int row0 = warp * 4 + (lane >> 3);       // this thread's row in pass 0; pass j adds j * ROWS_PER_PASS
    // 128B swizzle: chunk v of row r goes to chunk v ^ (r & 7); r & 7 is the same in every pass
    const uint32_t dst_off = (uint32_t)(row0 * 128 + ((v ^ (row0 & 7)) << 4));
    const uint32_t mp_addr = smem_u32(smem + L.meta_p) + (uint32_t)row0 * 16u;
    // byte steps to the right column / the lower row of an NHWC map (the same for every sample of this feature map)
    const uint32_t step_x = (uint32_t)(pr.x_sw * 2), step_y = (uint32_t)(pr.x_sh * 2);

    // ---- sample metadata: thread -> (row, CPT of its 4 corners) ----
    const int mrow = tid % ROWS, cg = tid / ROWS;
    const bool has_off = !PLAIN && !FCB && pr.offset != nullptr, has_mask = !PLAIN && !FCB && pr.mask != nullptr;
    const bool off_bf16 = (p.flags & FLAG_OFFSETS_BF16) != 0;      // internal flag: offsets/masks stored as bf16
    const bool mask_sig = has_mask && (p.flags & STM_DCN_MASK_SIGMOID);
    int hb = 0, wb = 0;                     // top-left of the un-deformed receptive field
    bool rvalid = false;
    int64_t off_base = 0, mask_base = 0;
    const __nv_bfloat16* ximg = reinterpret_cast<const __nv_bfloat16*>(pr.x);
    {
      const int m = m0 + mrow;
      if (m < pr.m_total) {
        const int hw = pr.out_h * pr.out_w;
        const int b = m / hw;
        const int r = m - b * hw;
        int ho, wo;
        row_to_pixel(pr, r, ho, wo);
        rvalid = true;
        hb = ho * p.sh - p.ph;
        wb = wo * p.sw - p.pw;
        ximg += (int64_t)b * pr.x_sn;
        off_base = b * pr.off_sn + ho * pr.off_sh + wo * pr.off_sw;
        mask_base = b * pr.mask_sn + ho * pr.mask_sh + wo * pr.mask_sw;
      }
    }
    const uint64_t zero_page = reinterpret_cast<uint64_t>(g_zero_page);

    // raw (dy, dx, mask) of this thread's row for iteration `it`; issued one iteration ahead of its use
    // (FCB: the four box deltas (t_x, t_y, t_w, t_h) of the row instead — the same for every tap, an L1 hit after the first)
    const bool fcb_ada = FCB && (p.flags & STM_DCN_FCB_ADA) != 0;
    auto load_raw = [&](int it_, float& oy, float& ox, float& mk, float& r3) {
      if (FCB && it_ > 0) return;              // the four box deltas are the same for every tap: loaded once, kept in registers
      oy = 0.f; ox = 0.f; mk = FCB ? 0.f : 1.f; r3 = 0.f;
      if (PLAIN || !rvalid || it_ >= n_iter) return;
      if (FCB) {
        if (off_bf16) {
          const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(pr.offset) + off_base;
          oy = __bfloat162float(__ldg(bp)); ox = __bfloat162float(__ldg(bp + pr.off_sc));
          mk = __bfloat162float(__ldg(bp + 2 * pr.off_sc)); r3 = __bfloat162float(__ldg(bp + 3 * pr.off_sc));
        } else {
          const float* bp = reinterpret_cast<const float*>(pr.offset) + off_base;
          oy = __ldg(bp); ox = __ldg(bp + pr.off_sc); mk = __ldg(bp + 2 * pr.off_sc); r3 = __ldg(bp + 3 * pr.off_sc);
        }
        return;
      }
      const int tap_ = it_ / p.dg, g_ = it_ - tap_ * p.dg;
      if (has_off) {
        const int64_t o = off_base + (int64_t)(g_ * 2 * K + 2 * tap_) * pr.off_sc;
        if (off_bf16) {
          oy = __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(pr.offset) + o));
          ox = __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(pr.offset) + o + pr.off_sc));
        } else {
          oy = __ldg(reinterpret_cast<const float*>(pr.offset) + o);
          ox = __ldg(reinterpret_cast<const float*>(pr.offset) + o + pr.off_sc);
        }
      }
      if (has_mask) {
        const int64_t o = mask_base + (int64_t)(g_ * K + tap_) * pr.mask_sc;
        mk = off_bf16 ? __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(pr.mask) + o))
                      : __ldg(reinterpret_cast<const float*>(pr.mask) + o);
      }
    };
    // DCN border rule: the sample is 0 outside (-1, H) x (-1, W); corners outside the map add 0.
    auto compute_meta = [&](int buf, int it_, float oy, float ox, float mk, float r3) {
      const int tap_ = it_ / p.dg, g_ = it_ - tap_ * p.dg;
      const int ti = tap_ / p.kw, tj = tap_ - ti * p.kw;
      if (FCB) {
        // (oy, ox, mk, r3) hold the box deltas (t_x, t_y, t_w, t_h)
        const float tx = oy, ty = ox, tw_ = mk, th_ = r3;
        if (fcb_ada) {
          const float4* fw = reinterpret_cast<const float4*>(smem + L.fcb_w) + (g_ * 2 * K + 2 * tap_);
          const float4 wy = fw[0], wx = fw[1];
          oy = fmaf(wy.w, th_, fmaf(wy.z, tw_, fmaf(wy.y, ty, wy.x * tx)));
          ox = fmaf(wx.w, th_, fmaf(wx.z, tw_, fmaf(wx.y, ty, wx.x * tx)));
        } else {
          // Featurealign.py:46-69: centre shift 0.1 * t * k, scale (exp(0.2 * t) - 1) * (tap index - k/2)
          oy = 0.1f * ty * (float)p.kh + (__expf(0.2f * th_) - 1.f) * (float)(ti - p.kh / 2);
          ox = 0.1f * tx * (float)p.kw + (__expf(0.2f * tw_) - 1.f) * (float)(tj - p.kw / 2);
        }
        mk = 1.f;
      }
      if (PLAIN) {
        // regular convolution: the tap's pixel itself, or the zero page when it lies in the padding
        if (cg == 0) {
          const int yy = hb + ti * p.dh, xx = wb + tj * p.dw;
          const bool ok = rvalid && yy >= 0 && yy < pr.in_h && xx >= 0 && xx < pr.in_w;
          const uint64_t ptr = ok ? reinterpret_cast<uint64_t>(ximg + ((int64_t)yy * pr.x_sh + (int64_t)xx * pr.x_sw + g_ * cpd)) : zero_page;
          *reinterpret_cast<uint2*>(smem + L.meta_p + (buf * ROWS + mrow) * 16) = make_uint2((uint32_t)ptr, (uint32_t)(ptr >> 32));
        }
        return;
      }
      // ---- compact 16-byte record per row: base pointer of the top-left corner (clamped into the map) with two flag
      //      bits in its low bits (bit 0: the right column is a different pixel, bit 1: the lower row is), and the
      //      four bf16 corner weights.  Corners that carry no weight (outside the map / outside the sample's support /
      //      rows past the end) keep weight 0 and read a clamped in-map address, so the gather stays branch-free
      //      and every load is in bounds; one LDS.128 per gather task instead of 40 bytes in three loads (the shared-
      //      memory / L1 data pipe is this kernel's busiest unit: profiles/r02_dcn_fcb35_pair2.txt). ----
      if (!CO && cg != 0) return;             // tap-major: group 0 builds the record; chunk-major: every group builds ONE tap's
      const float h = (float)(hb + ti * p.dh) + oy;
      const float w = (float)(wb + tj * p.dw) + ox;
      const bool inside = rvalid && h > -1.f && w > -1.f && h < (float)pr.in_h && w < (float)pr.in_w;
      const float hf = floorf(h), wf = floorf(w);
      const float lh = h - hf, lw = w - wf;
      // clamp before the float -> int conversion (huge offsets must not overflow the index arithmetic)
      const int h0 = (int)fminf(fmaxf(hf, -1.f), (float)pr.in_h), w0 = (int)fminf(fmaxf(wf, -1.f), (float)pr.in_w);
      const float scale = mask_sig ? sigmoidf_(mk) : mk;
      const bool y0ok = inside && h0 >= 0 && h0 < pr.in_h, y1ok = inside && h0 + 1 >= 0 && h0 + 1 < pr.in_h;
      const bool x0ok = inside && w0 >= 0 && w0 < pr.in_w, x1ok = inside && w0 + 1 >= 0 && w0 + 1 < pr.in_w;
      const int y0c = min(max(h0, 0), pr.in_h - 1), y1c = min(max(h0 + 1, 0), pr.in_h - 1);
      const int x0c = min(max(w0, 0), pr.in_w - 1), x1c = min(max(w0 + 1, 0), pr.in_w - 1);
      const uint32_t w00 = (y0ok && x0ok) ? (pack_bf16((1.f - lh) * (1.f - lw) * scale, 0.f) & 0xffffu) : 0u;
      const uint32_t w01 = (y0ok && x1ok) ? (pack_bf16((1.f - lh) * lw * scale, 0.f) & 0xffffu) : 0u;
      const uint32_t w10 = (y1ok && x0ok) ? (pack_bf16(lh * (1.f - lw) * scale, 0.f) & 0xffffu) : 0u;
      const uint32_t w11 = (y1ok && x1ok) ? (pack_bf16(lh * lw * scale, 0.f) & 0xffffu) : 0u;
      const uint64_t base = reinterpret_cast<uint64_t>(ximg + ((int64_t)y0c * pr.x_sh + (int64_t)x0c * pr.x_sw + g_ * cpd)) |
                            (uint64_t)((x1c != x0c ? 1 : 0) | (y1c != y0c ? 2 : 0));
      *reinterpret_cast<uint4*>(smem + L.meta_p + (buf * ROWS + mrow) * 16) =
          make_uint4((uint32_t)base, (uint32_t)(base >> 32), w00 | (w01 << 16), w10 | (w11 << 16));
    };

    // One gather task = (row, 8 channels) of one K block: 4 corner loads, fp32 blend (FHFMA.BF16: bf16 data and
    // weights feed the FMA directly), one rounding to bf16, one 16-byte store (PLAIN: one load, one store).  Every
    // thread keeps D tasks in flight while it blends and stores the current one.
    struct GTask {
      uint32_t w01, w23;   // bf16 corner weights (w0 | w1 << 16, w2 | w3 << 16)
      uint4 c[NC];
    };
    struct GMeta {
      uint4 r;             // base pointer lo / hi (flags in the low bits), w0 | w1 << 16, w2 | w3 << 16
    };
    auto read_meta = [&](GMeta& m, int buf, int j) {
      const uint32_t pa = mp_addr + (uint32_t)((buf * ROWS + j * ROWS_PER_PASS) * 16);
      if (PLAIN) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(m.r.x), "=r"(m.r.y) : "r"(pa));
        return;
      }
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(m.r.x), "=r"(m.r.y), "=r"(m.r.z), "=r"(m.r.w) : "r"(pa));
    };
    auto issue = [&](GTask& t, const GMeta& m, uint32_t coff) {
      if (PLAIN) {
        t.c[0] = ldg_nc_v4(add_wide(m.r.x, m.r.y, coff));
        return;
      }
      t.w01 = m.r.z; t.w23 = m.r.w;
      const uint32_t dx = (m.r.x & 1u) ? step_x : 0u, dy = (m.r.x & 2u) ? step_y : 0u;
      const uint64_t p00 = add_wide(m.r.x & ~3u, m.r.y, coff);
      const uint64_t p10 = p00 + dy;
      t.c[0] = ldg_nc_v4(p00);
      t.c[1 % NC] = ldg_nc_v4(p00 + dx);
      t.c[2 % NC] = ldg_nc_v4(p10);
      t.c[3 % NC] = ldg_nc_v4(p10 + dx);
    };
    auto finish = [&](const GTask& t, uint32_t dst) {
      uint32_t o[4];
      if (PLAIN) {
        o[0] = t.c[0].x; o[1] = t.c[0].y; o[2] = t.c[0].z; o[3] = t.c[0].w;
      } else {
        const uint32_t* q0 = reinterpret_cast<const uint32_t*>(&t.c[0]);
        const uint32_t* q1 = reinterpret_cast<const uint32_t*>(&t.c[1 % NC]);
        const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&t.c[2 % NC]);
        const uint32_t* q3 = reinterpret_cast<const uint32_t*>(&t.c[3 % NC]);
        uint16_t w0, w1, w2, w3;
        split16(t.w01, w0, w1);
        split16(t.w23, w2, w3);
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
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]));
    };

    STM_TR(1);                              // setup done (barriers, TMEM, cluster sync)
    GTask S[D];                             // task (kb, j) lives in S[j % D]; D tasks are in flight per thread
    float r_oy, r_ox, r_mk, r_3;
    // prologue: metadata of iteration 0, then the first D gather tasks of K block 0
    // Records are built ahead of their first use.  Tap-major: thread group 0 (one thread per row) builds tap it+1 at the first
    // chunk of tap it, three rotating buffers.  Chunk-major: record t lives in buffer t and is reused by every later chunk;
    // the NCG = PT / ROWS thread groups each build ONE tap's records per step (group g: tap it+1+g), so a step happens every
    // NCG K blocks of the first chunk pass — half the barriers and no idle half of the CTA (with 9 taps and 18 K blocks per
    // tile, the C = 128 stage had a record step with half its threads waiting in every second K block).
    constexpr int NCG = PT / ROWS;
    if (CO) {
      if (FCB) load_raw(0, r_oy, r_ox, r_mk, r_3);                    // the box deltas: the same for every tap
      else load_raw(cg, r_oy, r_ox, r_mk, r_3);
      if (cg < n_iter) compute_meta(cg, cg, r_oy, r_ox, r_mk, r_3);
      load_raw(NCG + cg, r_oy, r_ox, r_mk, r_3);
    } else {
      load_raw(0, r_oy, r_ox, r_mk, r_3);
      compute_meta(0, 0, r_oy, r_ox, r_mk, r_3);
      load_raw(1, r_oy, r_ox, r_mk, r_3);
    }
    named_barrier_sync(1, PT);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      GMeta m;
      read_meta(m, 0, j);
      issue(S[j], m, (uint32_t)(v * 16));
    }
    STM_TR(2);                              // first records computed, first loads issued
    int stage = 0;
    uint32_t phase = 0;
    int it = 0, cc = 0;                     // (tap, group) iteration and channel chunk of the CURRENT K block
#pragma unroll 1
    for (int kb = 0; kb < num_kb; ++kb) {
      if (CO) {
        if (cc == 0 && (it % NCG) == NCG - 1 && it + 1 < n_iter) {
          // chunk-major: records it+1 .. it+NCG, one per thread group, each buffer written once
          const int mine = it + 1 + cg;
          if (mine < n_iter) compute_meta(mine, mine, r_oy, r_ox, r_mk, r_3);
          load_raw(mine + NCG, r_oy, r_ox, r_mk, r_3);
          named_barrier_sync(1, PT);
        }
      } else if (cc == 0 && it + 1 < n_iter) {
        // tap-major, one iteration ahead: buffer (it+1) % 3 was last read by the gather of iteration it-2, which every thread
        // finished before it arrived at the previous barrier
        compute_meta((it + 1) % META_BUFS, it + 1, r_oy, r_ox, r_mk, r_3);
        load_raw(it + 2, r_oy, r_ox, r_mk, r_3);
        named_barrier_sync(1, PT);
      }
      // next K block: tap-major (tap, chunk) or, with resident records, chunk-major (chunk, tap) — consecutive K blocks
      // then read the SAME 128 bytes of neighbouring pixels, which is what lets L1 serve the taps' overlap
      int nit, ncc;
      if (CO) { nit = it + 1; ncc = cc; if (nit == K) { nit = 0; ncc = cc + 1; } }
      else { nit = it; ncc = cc + 1; if (ncc == chunks) { ncc = 0; nit = it + 1; } }
      const bool has_next = kb + 1 < num_kb;
      const int cbuf = CO ? it : it % META_BUFS, nbuf = CO ? nit : nit % META_BUFS;
      const uint32_t ccoff = (uint32_t)((cc * BLOCK_K + v * 8) * 2), ncoff = (uint32_t)((ncc * BLOCK_K + v * 8) * 2);
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      const uint32_t a_dst = smem_u32(smem + stage * L.stage_bytes) + dst_off;
      // step j: blend + store task (kb, j), then refill its register set with task j + D of this K block or,
      // past the end, task j + D - TPK of the next one
      GMeta m;
      if (D < TPK) read_meta(m, cbuf, D);
      else if (has_next) read_meta(m, nbuf, 0);
#pragma unroll
      for (int j = 0; j < TPK; ++j) {
        const bool cur = j + D < TPK;                                   // compile-time after unrolling
        const bool ncur = j + 1 + D < TPK;
        GMeta mn;
        if (j + 1 < TPK) {
          if (ncur) read_meta(mn, cbuf, j + 1 + D);
          else if (has_next) read_meta(mn, nbuf, j + 1 + D - TPK);
        }
        finish(S[j % D], a_dst + (uint32_t)(j * ROWS_PER_PASS * 128));
        if (cur) issue(S[j % D], m, ccoff);
        else if (has_next) issue(S[j % D], m, ncoff);
        if (j + 1 < TPK) m = mn;
      }
      // this warp's part of the A tile of `stage` is complete
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
      it = nit; cc = ncc;
    }
    // =============================== EPILOGUE ===============================
    STM_TR(3);                              // main loop done
    mbar_wait(accum_bar, 0);
    STM_TR(4);                              // accumulators complete
    tcgen05_fence_after();
    // warp w may only touch TMEM lanes [32 (w % 4), +32).  The PW / 4 warp groups split the accumulators:
    // M_TILES == 2 -> (tile, column part); M_TILES == 1 -> column part.
    constexpr int NG = PW / 4;
    const int q = warp & 3;
    const int grp = warp >> 2;
    const int mt = (M_TILES == 2) ? (grp & 1) : 0;
    const int part = (M_TILES == 2) ? (grp >> 1) : grp;
    constexpr int PARTS = (M_TILES == 2) ? NG / 2 : NG;
    const int nchunk = block_n / 16;
    const int c_begin = (part * nchunk / PARTS) * 16;
    const int c_end = ((part + 1) * nchunk / PARTS) * 16;
    const int row = mt * TILE_M + q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < pr.m_total;
    int64_t yoff = 0;
    if (row_ok) {
      const int hw = pr.out_h * pr.out_w;
      const int b = m / hw;
      const int r = m - b * hw;
      int ho, wo;
      row_to_pixel(pr, r, ho, wo);
      yoff = b * pr.y_sn + ho * pr.y_sh + wo * pr.y_sw;
    }
    const bool relu = (p.flags & STM_DCN_RELU) != 0;
    const bool out_f32 = (p.flags & STM_DCN_OUT_F32) != 0;
    const bool planar = (p.flags & STM_DCN_OUT_PLANAR) != 0;       // y[b][n][ho][wo]: consecutive lanes = consecutive pixels
    const int64_t y_sc = (int64_t)pr.out_h * pr.y_sh;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * block_n);
    // bf16 NHWC output: stage the tile in shared memory (the operand ring and the sample records are dead once the accumulators
    // are complete) and write it out with every warp instruction covering ONE row's contiguous bytes.  A thread owns a TMEM
    // lane = an output row, so storing straight from registers makes each 16-byte STG of a warp hit 32 different lines: 32 L1
    // wavefronts instead of 4 per instruction, 6-11 k clk of a CTA's 58-150 k (STM_DCN_TRACE).
    const int pitch = block_n * 2 + 16;                                  // +16 B: lanes 8 rows apart share a bank group, 4 wavefronts per STS.128
    const bool staged = !planar && !out_f32 && ROWS * pitch + ROWS * 8 <= L.bars;
    if (staged) {
      uint8_t* stag = smem;
      int64_t* s_yoff = reinterpret_cast<int64_t*>(smem + ROWS * pitch);
      if (part == 0) s_yoff[row] = row_ok ? yoff : -1;
      auto convert_store = [&](const uint32_t* acc, int c0) {           // 16 columns -> bias, ReLU, bf16 -> 32 bytes of the staged row
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          f[i] = __uint_as_float(acc[i]);
          if (p.bias != nullptr) f[i] += __ldg(p.bias + n0 + c0 + i);
          if (relu) f[i] = fmaxf(f[i], 0.f);
        }
        uint4* dst = reinterpret_cast<uint4*>(stag + row * pitch + c0 * 2);
        dst[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        dst[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
      };
      int c0 = c_begin;
      for (; c0 + 32 <= c_end; c0 += 32) {                             // 32 columns per TMEM round trip
        uint32_t acc[32];
        tmem_ld32(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        convert_store(acc, c0);
        convert_store(acc + 16, c0 + 16);
      }
      for (; c0 < c_end; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        convert_store(acc, c0);
      }
      named_barrier_sync(1, PT);
      const int cpr = block_n / 8;                                       // 16-byte chunks per row
      __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(pr.y) + n0;
      if (cpr <= 32 && (cpr & (cpr - 1)) == 0) {
        // a warp instruction = 32 / cpr whole rows (N = 256: one row, 512 contiguous bytes); shifts, no divisions
        const int sh = __ffs(cpr) - 1;
        const int rsub = lane >> sh, ch = lane & (cpr - 1), rpi = 32 >> sh;
        for (int r = warp * rpi + rsub; r < ROWS; r += PW * rpi) {
          const int64_t yo = s_yoff[r];
          if (yo >= 0) *reinterpret_cast<uint4*>(yb + yo + ch * 8) = *reinterpret_cast<const uint4*>(stag + r * pitch + ch * 16);
        }
      } else {
        for (int i = tid; i < ROWS * cpr; i += PT) {
          const int r = i / cpr, ch = i - r * cpr;
          const int64_t yo = s_yoff[r];
          if (yo >= 0) *reinterpret_cast<uint4*>(yb + yo + ch * 8) = *reinterpret_cast<const uint4*>(stag + r * pitch + ch * 16);
        }
      }
    } else {
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      uint32_t acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (row_ok) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          f[i] = __uint_as_float(acc[i]);
          if (p.bias != nullptr) f[i] += __ldg(p.bias + n0 + c0 + i);
          if (relu) f[i] = fmaxf(f[i], 0.f);
        }
        if (planar) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int64_t o = yoff + (int64_t)(n0 + c0 + i) * y_sc;
            if (out_f32) reinterpret_cast<float*>(pr.y)[o] = f[i];
            else reinterpret_cast<__nv_bfloat16*>(pr.y)[o] = __float2bfloat16_rn(f[i]);
          }
        } else if (out_f32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(pr.y) + yoff + n0 + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pr.y) + yoff + n0 + c0);
          dst[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          dst[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
        }
      }
    }
    }
#ifdef STM_DCN_TRACE
    STM_TR(5);
    if (tid == 0 && (blockIdx.x == 1000 || blockIdx.x == 1001 || blockIdx.x == 1600))
      printf("dcn_tc cta %d: setup %lld, first records+loads %lld, main loop %lld (%d K blocks), wait accum %lld, epilogue %lld\n", (int)blockIdx.x,
             tr_[1] - tr_[0], tr_[2] - tr_[1], tr_[3] - tr_[2], num_kb, tr_[4] - tr_[3], tr_[5] - tr_[4]);
#endif
  } else if (warp == PW) {
    // =============================== TMA (weights) ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      // a pair splits the N tile: this CTA loads weight rows [n0 + rank * block_n/2, + block_n/2)
      const int nrow0 = PAIR ? n0 + (int)cta_rank * (block_n / 2) : n0;
      const int outer = CO ? chunks : K, inner = CO ? K : chunks;       // CO: dg == 1
#pragma unroll 1
      for (int o = 0; o < outer; ++o)
#pragma unroll 1
        for (int g = 0; g < p.dg; ++g)
#pragma unroll 1
          for (int i = 0; i < inner; ++i) {
            const int tap = CO ? i : o, cc = CO ? o : i;
            mbar_wait_relaxed(&empty_bar[s], phase ^ 1u);
            uint8_t* dst = smem + s * L.stage_bytes + M_TILES * A_TILE_BYTES;
            const int kcol = tap * p.in_c + g * cpd + cc * BLOCK_K;
            if (PAIR) {
              if (leader) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(block_n * 128));     // both halves
              tma_load_2d_pair(dst, &tmap_w, full_leader + (uint32_t)s * 8u, kcol, nrow0);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(block_n * 128));
              tma_load_2d(dst, &tmap_w, &full_bar[s], kcol, nrow0);
            }
            if (++s == stages) { s = 0; phase ^= 1u; }
          }
    }
  } else {
    // =============================== MMA issuer (a pair: the leader CTA only) ===============================
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * TILE_M : TILE_M, (uint32_t)block_n);
      int s = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        if (a.pad_ & 1) mbar_wait(&full_bar[s], phase); else mbar_wait_relaxed(&full_bar[s], phase);
        fence_proxy_async_smem();              // this CTA's producers' st.shared (acquired above) -> visible to the UMMA reads
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L.stage_bytes);
        const uint64_t bdesc = umma_desc_sw128(a_addr + M_TILES * A_TILE_BYTES);
#pragma unroll
        for (int mt = 0; mt < M_TILES; ++mt) {
          const uint64_t adesc = umma_desc_sw128(a_addr + mt * A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            if (PAIR) umma_bf16_pair(tmem_base + (uint32_t)(mt * block_n), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
            else umma_bf16(tmem_base + (uint32_t)(mt * block_n), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
          }
        }
        // stage reusable (in both CTAs of a pair) once these MMAs have read it
        if (PAIR) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
        if (++s == stages) { s = 0; phase ^= 1u; }
      }
      if (PAIR) umma_commit_pair(accum_bar); else umma_commit(accum_bar);     // accumulators complete
    } else if (PAIR && lane == 0) {
      // ---- relay (non-leader CTA of a pair): A tile complete -> proxy fence on THIS SM -> one arrive on the leader ----
      int s = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        if (a.pad_ & 1) mbar_wait(&full_bar[s], phase); else mbar_wait_relaxed(&full_bar[s], phase);
        fence_proxy_async_smem();
        mbar_arrive_remote(full_leader + (uint32_t)s * 8u);
        if (++s == stages) { s = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  }

  // ---- teardown ----
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == PW + 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)a.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int pick_block_n(int out_c) {
  if (out_c <= 256) return out_c;
  if (out_c % 256 == 0) return 256;
  if (out_c % 192 == 0) return 192;
  if (out_c % 128 == 0) return 128;
  return 0;
}

SmemAttrCache g_smem_attr[24];         // one per instantiation below

template <int M_TILES, int PW, int D, bool PAIR, int MODE>
int launch_t(const TcArgs& args, const CUtensorMap& tmap, dim3 grid, int smem_bytes, cudaStream_t stream) {
  constexpr int slot = (((M_TILES - 1) * 2 + (PW == 16 ? 1 : 0)) * 2 + (PAIR ? 1 : 0)) * 3 + MODE;
  auto kernel = dcn_tc_kernel<M_TILES, PW, D, PAIR, MODE>;
  const int rc = ensure_dynamic_smem(kernel, smem_bytes, g_smem_attr[slot]);
  if (rc != STM_OK) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(PW * 32 + 64);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.numAttrs = 0;
  if (PAIR) {
    attr[0].id = cudaLaunchAttributeClusterDimension;      // the two CTAs of a pair: consecutive blockIdx.x, same TPC
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  STM_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, args, tmap));
  count_launch();
  return STM_OK;
}

// Everything the host decides about a launch: CTA shape, pairing, pipeline depth, grid.  A pure function of the
// call's arguments and the device's SM count (no environment variables, no mutable state), shared by the
// launcher and by stm_deform_conv2d_variant().
struct TcPlan {
  int block_n, n_tiles, m_tiles, pw, stages, smem_bytes, tmem_cols, blocks, grid_x, meta_bufs;
  bool two_ctas, pair, plain, fcb, chunk_outer;
};

int make_plan(const StmDcnConv* conv, const DcnParams& p, TcPlan* out) {
  TcPlan pl;
  pl.block_n = pick_block_n(p.out_c);
  pl.plain = (p.flags & STM_DCN_ZERO_OFFSET) != 0;
  pl.fcb = !pl.plain && (p.flags & (STM_DCN_FCB_ADA | STM_DCN_FCB_ALI)) != 0;
  const int fcb_floats = pl.fcb ? p.dg * 2 * p.kh * p.kw * 4 : 0;
  int64_t rows = 0;
  for (int i = 0; i < p.n_probs; ++i) rows += p.prob[i].m_total;
  pl.n_tiles = p.out_c / pl.block_n;
  // two accumulators (256 rows) per CTA halve the weight traffic per row (B200: 0.205 -> 0.150 ms on the 24x40
  // backbone layers); fall back to 128-row CTAs only when 256-row CTAs could not even half-fill the GPU
  const int64_t ctas256 = ((rows + 255) / 256) * pl.n_tiles;
  pl.m_tiles = (ctas256 >= device_sm_count() / 2 && 2 * pl.block_n <= 512) ? 2 : 1;
  // N <= 128 (the C = 128 backbone stage): the MMA is cheap, the gather and the per-CTA prologue / epilogue
  // dominate.  Two 128-row CTAs with 8 producer warps each share an SM (registers, 128 TMEM columns and
  // < 113 KB shared memory each), so one CTA's prologue / epilogue hides behind the other's main loop
  // (B200: 0.358 -> 0.323 ms on the 48x80 C=128 layers).  With N = 256 the doubled weight traffic costs more.
  pl.two_ctas = pl.block_n <= 128 && pl.block_n >= 64 && !pl.plain;
  // N = 256 with enough rows to put two 128-row CTAs on every SM: two 8-warp CTAs per SM, each the half of a
  // cta_group::2 PAIR with a CTA on the neighbouring SM (so it keeps only 128 of the 256 weight rows: 32 KB per
  // stage, 80 KB per CTA, 256 TMEM columns).  One CTA's prologue / epilogue and the pair's lock-step stalls hide
  // behind the other CTA's main loop (B200, FCB 3x5 over P3..P7, 72 frames: 1.008 -> 0.965 ms; 256-row pairs
  // with the same two stages: 1.055 ms, with three stages and less L1: 1.005 ms).
  const int64_t ctas128 = ((rows + 127) / 128) * pl.n_tiles;
  bool pair2 = pl.block_n == 256 && !pl.plain && ctas128 >= 2 * device_sm_count();
  if (pl.two_ctas) pl.m_tiles = 1;
  // explicit, stateless overrides (tests force every CTA shape through the same entry point)
  if ((conv->flags & (STM_DCN_HINT_ROWS128 | STM_DCN_HINT_ROWS256 | STM_DCN_HINT_NO_PAIR)) != 0) pair2 = false;
  if ((conv->flags & STM_DCN_HINT_ROWS128) != 0) { pl.m_tiles = 1; pl.two_ctas = false; }
  if ((conv->flags & STM_DCN_HINT_ROWS256) != 0 && 2 * pl.block_n <= 512) { pl.m_tiles = 2; pl.two_ctas = false; }
  if ((conv->flags & STM_DCN_HINT_TWO_CTAS) != 0 && !pl.plain && pl.block_n >= 64) {
    pl.m_tiles = 1;
    pl.two_ctas = true;
    pair2 = pl.block_n >= 128 && (conv->flags & STM_DCN_HINT_NO_PAIR) == 0;
  }
  if (pair2) { pl.m_tiles = 1; pl.two_ctas = true; }
  pl.pw = pl.two_ctas ? 8 : 16;
  // CTA pairs (tcgen05 cta_group::2): each CTA fetches and keeps only half of the N tile's weights.  On their own
  // (one 256-row CTA per SM) they measured no faster than unpaired CTAs — the kernel is bound by gather latency,
  // not by shared-memory bandwidth, although the pair does cut the tensor core's shared-memory reads by a third
  // (profiles/r02_dcn_pair_vs_single.txt) — so they are used where they let two CTAs share an SM, and for the
  // plain-conv mode with wide N; STM_DCN_HINT_ROWS256 / _ROWS128 without _NO_PAIR still select them for the tests.
  const bool hinted_rows = (conv->flags & (STM_DCN_HINT_ROWS128 | STM_DCN_HINT_ROWS256)) != 0;
  pl.pair = (pair2 || (!pl.two_ctas && (hinted_rows || pl.plain) && pl.block_n >= 128)) && pl.block_n % 32 == 0 &&
            (conv->flags & STM_DCN_HINT_NO_PAIR) == 0;
  const int rows_per_cta = TILE_M * pl.m_tiles;
  pl.blocks = 0;
  for (int i = 0; i < p.n_probs; ++i) pl.blocks += (p.prob[i].m_total + rows_per_cta - 1) / rows_per_cta;
  if (pl.pair && pl.blocks * pl.n_tiles < device_sm_count()) pl.pair = false;      // small launch: keep every tile on its own SM
  pl.grid_x = pl.pair ? (pl.blocks + 1) & ~1 : pl.blocks;                           // a pair's odd filler CTA owns no rows
  int cols = 32;
  while (cols < pl.m_tiles * pl.block_n) cols <<= 1;
  pl.tmem_cols = cols;
  // pipeline depth: as many stages as fit while leaving L1 some room for the gather's corner reuse
  int budget = pl.m_tiles == 2 ? 172 * 1024 : (pl.two_ctas ? (pl.pair ? 82 * 1024 : 110 * 1024) : 132 * 1024);
  budget += (fcb_floats * 4 + 15) & ~15;
  if ((conv->flags & STM_DCN_HINT_DEEP_PIPE) != 0) budget = 200 * 1024;
  // chunk-major K order: all kh*kw sample records of a row stay in shared memory (16 B each), so consecutive K blocks are
  // the neighbouring taps of ONE 64-channel chunk and re-read each other's cache lines from L1 instead of L2
  // (B200, 1024 frames, with the 8x8 patch row order: FCB 3x5 13.69 -> 13.11 ms, C = 512 s2 2.08 -> 1.58 ms, C = 128 s2
  // 4.18 -> 3.76 ms with two stages instead of three so that two CTAs still share an SM).  The default whenever
  // deform_groups == 1 and there is more than one chunk per tap; STM_DCN_HINT_TAP_MAJOR keeps the other order.
  pl.chunk_outer = !pl.plain && p.dg == 1 && p.in_c > BLOCK_K && p.kh * p.kw > 1 && p.kh * p.kw <= 25 &&
                   (conv->flags & STM_DCN_HINT_TAP_MAJOR) == 0;
  pl.meta_bufs = pl.chunk_outer ? p.kh * p.kw : META_BUFS;
  if (pl.chunk_outer) budget += (pl.meta_bufs - META_BUFS) * rows_per_cta * 16;
  if (pl.two_ctas && budget > 113 * 1024) budget = 113 * 1024;          // both CTAs must still fit one SM's 227 KB
  auto total = [&](int st) {
    return pl.m_tiles == 2 ? SmemLayout<2>(pl.block_n, st, pl.pair, fcb_floats, pl.meta_bufs).total
                           : SmemLayout<1>(pl.block_n, st, pl.pair, fcb_floats, pl.meta_bufs).total;
  };
  pl.stages = MAX_STAGES;
  for (; pl.stages > 2; --pl.stages)
    if (total(pl.stages) <= budget) break;
  pl.smem_bytes = total(pl.stages);
  if (pl.smem_bytes > SMEM_LIMIT) { set_error("tcgen05 DCN: shared memory %d B over the limit", pl.smem_bytes); return STM_ERR_UNSUPPORTED; }
  *out = pl;
  return STM_OK;
}

}  // namespace

bool dcn_tc_supported(const StmDcnConv* c, const StmDcnProblem* pr, int n, const char** why) {
  if (!dcn_tc_shape_supported(c, pr, n, why)) return false;
  if (get_tensormap_encoder() == nullptr) { *why = "cuTensorMapEncodeTiled unavailable"; return false; }
  return true;
}

// everything but the driver: also answers on a machine without a GPU (launch-plan queries)
bool dcn_tc_shape_supported(const StmDcnConv* c, const StmDcnProblem* pr, int n, const char** why) {
  *why = "";
  if (c->dtype != STM_BF16) { *why = "dtype is not bf16"; return false; }
  if (c->groups != 1) { *why = "groups != 1"; return false; }
  if (2 * c->in_c > ZERO_PAGE_BYTES) { *why = "in_c > 2048"; return false; }
  if (c->in_c % 64 != 0 || (c->in_c / c->deform_groups) % 64 != 0) { *why = "channels per deformable group not a multiple of 64"; return false; }
  if (c->out_c % 16 != 0 || pick_block_n(c->out_c) == 0) { *why = "out_c not tileable (multiple of 16, <= 256 or a multiple of 128)"; return false; }
  if (c->kernel_h * c->kernel_w > 64) { *why = "kernel too large"; return false; }
  for (int i = 0; i < n; ++i) {
    const StmDcnProblem& q = pr[i];
    if (q.batch == 0) continue;
    if (((uintptr_t)q.x & 15) || (!(c->flags & STM_DCN_OUT_PLANAR) && ((uintptr_t)q.y & 15))) { *why = "x / y not 16-byte aligned"; return false; }
    const bool planar = (c->flags & STM_DCN_OUT_PLANAR) != 0;      // scalar stores: no alignment demand on y
    if (((q.x_stride_n | q.x_stride_h | q.x_stride_w) & 7) || (!planar && ((q.y_stride_n | q.y_stride_h | q.y_stride_w) & 7))) {
      *why = "x / y strides not multiples of 8 elements";
      return false;
    }
  }
  return true;
}

size_t dcn_tc_workspace(const StmDcnConv*, const StmDcnProblem*, int) { return 0; }

int dcn_tc_variant(const StmDcnConv* conv, const DcnParams& p, char* buf, size_t len) {
  const char* why = "";
  if (conv_tma_shape_supported(conv, p, &why)) {
    const int rc = conv_tma_variant(conv, p, buf, len);
    if (rc != STM_ERR_UNSUPPORTED) return rc;
    clear_error();
  }
  TcPlan pl;
  const int rc = make_plan(conv, p, &pl);
  if (rc != STM_OK) return rc;
  snprintf(buf, len, "tcgen05 rows=%d n=%d pair=%d plain=%d fcb=%d producer_warps=%d stages=%d ctas_per_sm=%d korder=%s grid=%dx%d", TILE_M * pl.m_tiles,
           pl.block_n, pl.pair ? 1 : 0, pl.plain ? 1 : 0, pl.fcb ? 1 : 0, pl.pw, pl.stages, pl.two_ctas ? 2 : 1, pl.chunk_outer ? "chunk" : "tap",
           pl.grid_x, pl.n_tiles);
  return STM_OK;
}

static bool pl_fcb_needs_weight(const DcnParams& p) { return (p.flags & STM_DCN_FCB_ADA) != 0 && (p.flags & STM_DCN_ZERO_OFFSET) == 0; }

int launch_dcn_tc(const StmDcnConv* conv, const DcnParams& p_in, void*, size_t, cudaStream_t stream) {
  TcArgs args;
  args.p = p_in;
  DcnParams& p = args.p;
  p.flags &= 0x2ffff;                          // public flags (bit 16 is internal)
  if (pl_fcb_needs_weight(p) && p.fcb_w == nullptr) { set_error("FCB(ada) needs the conv_offset weight"); return STM_ERR_INVALID_ARGUMENT; }
  if (conv->offset_dtype == STM_BF16) p.flags |= FLAG_OFFSETS_BF16;
  {
    // regular convolutions that tile well: A operand by TMA, taps as shifted descriptor views (no gather at all)
    const char* why = "";
    if (conv_tma_shape_supported(conv, p_in, &why)) {
      const int rc = launch_conv_tma(conv, p_in, stream);
      if (rc != STM_ERR_UNSUPPORTED) return rc;            // a tensor map the driver refuses (exotic strides): the gather loop below takes it
      clear_error();
    }
  }
  TcPlan pl;
  const int prc = make_plan(conv, p, &pl);
  if (prc != STM_OK) return prc;
  const int rows_per_cta = TILE_M * pl.m_tiles;
  int blocks = 0;
  for (int i = 0; i < p.n_probs; ++i) {
    p.prob[i].tile_begin = blocks;
    blocks += (p.prob[i].m_total + rows_per_cta - 1) / rows_per_cta;
    p.prob[i].patch = (!pl.plain && (conv->flags & STM_DCN_HINT_RASTER) == 0 && p.prob[i].out_h % 8 == 0 && p.prob[i].out_w % 8 == 0) ? 1 : 0;
  }
  p.total_m_tiles = blocks;
  if (blocks == 0) return STM_OK;
  args.block_n = pl.block_n;
  args.tmem_cols = pl.tmem_cols;
  args.pad_ = (conv->flags >> 20) & 0xff;      // experiment bits (profiling only; undocumented, results identical)
  args.stages = pl.stages;
  args.chunk_outer = pl.chunk_outer ? 1 : 0;
  args.pad2_ = 0;

  CUtensorMap tmap;
  const cuuint64_t ktot = (cuuint64_t)p.kh * p.kw * p.in_c;
  const cuuint64_t dims[2] = {ktot, (cuuint64_t)p.out_c};
  const cuuint64_t strides[1] = {ktot * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(pl.pair ? pl.block_n / 2 : pl.block_n)};
  const cuuint32_t estr[2] = {1, 1};
  PFN_stm_encodeTiled enc = get_tensormap_encoder();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled unavailable"); return STM_ERR_CUDA; }
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p.w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return STM_ERR_CUDA; }

  const dim3 grid((unsigned)pl.grid_x, (unsigned)pl.n_tiles);
#define STM_GO(MT, PW_, D_, PAIR_, MODE_) return launch_t<MT, PW_, D_, PAIR_, MODE_>(args, tmap, grid, pl.smem_bytes, stream)
  if (pl.plain) {                      // regular convolution: every task of the next K block in flight (4 registers each)
    if (pl.m_tiles == 2) { if (pl.pair) STM_GO(2, 16, 4, true, 1); STM_GO(2, 16, 4, false, 1); }
    if (pl.pair) STM_GO(1, 16, 2, true, 1);
    STM_GO(1, 16, 2, false, 1);
  }
  if (pl.fcb) {                        // box-guided offsets derived in the metadata stage
    if (pl.m_tiles == 2) { if (pl.pair) STM_GO(2, 16, 2, true, 2); STM_GO(2, 16, 2, false, 2); }
    if (pl.pw == 8) { if (pl.pair) STM_GO(1, 8, 2, true, 2); STM_GO(1, 8, 2, false, 2); }
    if (pl.pair) STM_GO(1, 16, 2, true, 2);
    STM_GO(1, 16, 2, false, 2);
  }
  if (pl.m_tiles == 2) { if (pl.pair) STM_GO(2, 16, 2, true, 0); STM_GO(2, 16, 2, false, 0); }
  if (pl.pw == 8) { if (pl.pair) STM_GO(1, 8, 2, true, 0); STM_GO(1, 8, 2, false, 0); }   // 4 gather tasks per thread per K block, two CTAs per SM
  if (pl.pair) STM_GO(1, 16, 2, true, 0);
  STM_GO(1, 16, 2, false, 0);                           // 2 tasks
#undef STM_GO
}

}  // namespace stm
