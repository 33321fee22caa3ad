// Correlation cost volume on tcgen05 tensor cores.
//
// out[b, ph*P+pw, y, x] = post(scale * <x1[b,y,x,:], x2[b, y+(ph-r)d, x+(pw-r)d, :]>)
//
// The C-contraction is a GEMM between a PATCH of x1 pixels and the HALO REGION of x2 pixels around it:
//   tile    = TH x TW x1 pixels  (<= 128 = M rows,    K-major [pixel][channel])
//   region  = (TH+P-1) x (TW+P-1) x2 pixels (N <= 512, K-major [pixel][channel])
//   D[128, N] = X1_tile * X2_region^T, fp32 in TMEM; the P*P wanted displacements of pixel (py, px) are
//   the columns (py+ph) * RW + (px+pw)  — a lane-dependent band that the epilogue picks out.
// The tile shape is chosen per feature-map size (8x16, or 6x20 which tiles a 24x40 map exactly).
// Both operands arrive by 4-D TMA straight from the NHWC tensors (box {64 ch, w, h, 1}, 128B swizzle);
// out-of-map halo pixels are zero-filled by TMA, which IS the operator's border rule.  A dilated patch
// (dilation_patch = d) is d*d independent undilated problems on the sub-lattices (y % d, x % d); TMA
// element strides {1, d, d, 1} load a sub-lattice directly, so the same kernel serves every d.
//
// Per persistent CTA (one per SM):  warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-17 epilogue:
//   phase 0  (independent of the accumulator => overlaps the MMA) concat copy of the two T2S feature maps,
//   phase 1  tcgen05.ld -> scale / leaky-ReLU / ReLU -> band scatter into a shared staging tile [pixel][P*P],
//   phase 2  staging -> global in the caller's layout.
// FAST epilogue: when the output is channels-last bf16 and the feature block starts at a channel offset that
// is a multiple of 8 (the padded concat layout [corr P*P | zeros | feat_a | feat_b], e.g. 121 -> 128), every
// pixel row is 16-byte aligned: phase 0 is LDG.128 -> max.bf16x2 -> STG.128, phase 2 is LDS.64 x2 -> STG.128.
// GENERIC epilogue: any strides / dtypes, element-wise.
//
// Replaces correlation_cuda_forward_kernel (+ its two NHWC permute copies) of
// spatial_correlation_sampler and the elementwise tail of correlate()/CandidateShift
// (reference track_to_segment_head.py:53-62, TF_utils.py:28-31).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace stm {
namespace {

using namespace tc;

constexpr int STAGES = 2;
constexpr int DRAIN_WARPS = 8;                  // TMEM -> staging -> global (phases 1 and 2)
constexpr int COPY_WARPS = 8;                   // concat copy of the feature maps (phase 0), free-running
constexpr int DRAIN_THREADS = DRAIN_WARPS * 32;
constexpr int NUM_THREADS = 64 + DRAIN_THREADS + COPY_WARPS * 32;
constexpr int MAX_P = 11;

constexpr int MAX_LEVELS = 8;              // feature maps per grouped launch (the five FPN levels of the operator sweep)

// One feature map of a grouped launch: every level has its own size, output and tensor maps; C, patch, dilation,
// dtypes and post-ops are shared (args.d).  A single-map call is a grouped launch with one level.
struct CorrLevel {
  int32_t h, w, tile_begin, tiles_per_image;
  void* out;
  int64_t out_stride_n, out_stride_c, out_stride_h, out_stride_w;
};

struct CorrMaps {
  CUtensorMap x1[MAX_LEVELS], x2[MAX_LEVELS], x1_alt;
};

struct CorrTcArgs {
  StmCorrDesc d;
  CorrLevel lv[MAX_LEVELS];
  int32_t n_levels, pad0_;
  const void* fa;
  const void* fb;
  int32_t n_tiles, pad1_;
  int32_t th;              // tile height (tile width is the template parameter); th * TW <= 128
  int32_t rh, rw;          // region height / width in (sub-lattice) pixels
  int32_t n_half;          // N of each of the two MMAs (multiple of 16, <= 256)
  int32_t chunks;          // C / 64
  int32_t tmem_cols;
  int32_t stage_stride;    // elements per pixel in the staging tile
  int32_t fast;            // aligned channels-last epilogue
  int32_t feat_off;        // first feature channel (>= P*P)
  int32_t l2_prefetch;
  unsigned long long* trace;   // optional globaltimer trace (STM_DEBUG_BUF): 64 slots per CTA, see tools/corr_trace.py
};

struct SmemPlan {
  int a_bytes, b_bytes, stage_bytes, staging, bars, total;
  __host__ __device__ SmemPlan(int n_half, int stage_stride, int out_esize) {
    a_bytes = 128 * 128;
    b_bytes = 2 * n_half * 128;
    stage_bytes = a_bytes + b_bytes;
    staging = STAGES * stage_bytes;
    bars = staging + ((128 * stage_stride * out_esize + 15) & ~15);
    total = bars + 128 + 1024;
  }
};

struct TileCoord {
  int b, sy, sx, y0, x0;   // batch, sub-lattice phase, first pixel of the patch IN SUB-LATTICE coordinates
};

__device__ __forceinline__ int find_level(const CorrTcArgs& a, int tile) {
  int l = 0;
#pragma unroll 1
  for (int i = 1; i < a.n_levels; ++i)
    if (tile >= a.lv[i].tile_begin) l = i;
  return l;
}

// `tile` is relative to the level's first tile
__device__ __forceinline__ TileCoord decode_tile(const StmCorrDesc& dd, const CorrLevel& d, int tile, int th, int tw) {
  TileCoord t;
  const int tiles_per_image = d.tiles_per_image;
  t.b = tile / tiles_per_image;
  int rem = tile - t.b * tiles_per_image;
  const int dl = dd.dilation_patch;
  t.sy = t.sx = t.y0 = t.x0 = 0;
  for (int sy = 0; sy < dl; ++sy)
    for (int sx = 0; sx < dl; ++sx) {
      const int hs = (d.h - sy + dl - 1) / dl, ws = (d.w - sx + dl - 1) / dl;
      const int ty = (hs + th - 1) / th, tx = (ws + tw - 1) / tw;
      const int n = ty * tx;
      if (rem >= 0 && rem < n) {
        t.sy = sy; t.sx = sx;
        t.y0 = (rem / tx) * th;
        t.x0 = (rem % tx) * tw;
        rem = -1;
      } else if (rem >= 0) {
        rem -= n;
      }
    }
  return t;
}

// frames of pair `pair`: b1 indexes x1 / feat_a (or, when alt, x1_alt / feat_a_alt), b2 indexes x2 / feat_b
struct PairFrames {
  int b1, b2;
  bool alt;
};
__device__ __forceinline__ PairFrames pair_frames(const StmCorrDesc& d, int pair) {
  PairFrames f;
  f.b1 = f.b2 = pair;
  f.alt = false;
  if (d.x1_index != nullptr) {
    f.b1 = __ldg(d.x1_index + pair);
    f.b2 = __ldg(d.x2_index + pair);
    if (f.b1 >= d.x1_frames) { f.b1 -= d.x1_frames; f.alt = true; }
  }
  return f;
}

__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0u));
  return r;
}

// POST: 0 = scale only, 1 = leaky-ReLU, 2 = ReLU (ReLU after leaky-ReLU is ReLU)
template <typename OT, int TW, int POST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_tc_kernel(const __grid_constant__ CorrTcArgs a, const __grid_constant__ CorrMaps maps) {
  constexpr int MAX_RW = TW + MAX_P - 1;          // <= 32: one tcgen05.ld.x32 covers a region row
  static_assert(MAX_RW <= 32, "tile too wide");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemPlan L(a.n_half, a.stage_stride, (int)sizeof(OT));
  OT* staging = reinterpret_cast<OT*>(smem + L.staging);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const StmCorrDesc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = d.patch, PP = P * P, r = P / 2, dl = d.dilation_patch;
  const int th = a.th, npix = th * TW;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.x1[0]);
    prefetch_tensormap(&maps.x2[0]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, DRAIN_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  if (a.fast && warp >= 2) {
    // the pad channels [P*P, feat_off) of the staging rows stay zero for the whole kernel
    uint32_t* z = reinterpret_cast<uint32_t*>(staging);
    for (int i = tid - 64; i < (128 * a.stage_stride * (int)sizeof(OT)) / 4; i += NUM_THREADS - 64) z[i] = 0u;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long* trace = a.trace ? a.trace + (size_t)blockIdx.x * 64 : nullptr;
#define STM_TRACE(k_, e_)                                                             \
  do {                                                                                \
    if (trace && (k_) < 5) {                                                          \
      unsigned long long t_;                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
      trace[(k_) * 12 + (e_)] = t_;                                                   \
    }                                                                                 \
  } while (0)
  if (tid == 0) STM_TRACE(4, 11);      // slot 59: kernel body starts

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int s = 0, k = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)((npix + a.rh * a.rw) * 128);
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
        const int lvl = find_level(a, tile);
        const TileCoord t = decode_tile(d, a.lv[lvl], tile - a.lv[lvl].tile_begin, th, TW);
        const int fy = t.sy + dl * t.y0, fx = t.sx + dl * t.x0;           // full-resolution coords of the patch origin
        const PairFrames pf = pair_frames(d, t.b);
        const CUtensorMap* m1 = pf.alt ? &maps.x1_alt : &maps.x1[lvl];
        const CUtensorMap* m2 = &maps.x2[lvl];
        for (int c = 0; c < a.chunks; ++c) {
          mbar_wait_relaxed(&empty_bar[s], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          uint8_t* st = smem + s * L.stage_bytes;
          tma_load_4d(st, m1, &full_bar[s], c * 64, fx, fy, pf.b1);
          tma_load_4d(st + L.a_bytes, m2, &full_bar[s], c * 64, fx - r * dl, fy - r * dl, pf.b2);
          if (a.l2_prefetch) {
            // the smem ring holds only STAGES chunks: pull the chunk that will be loaded STAGES steps from now into
            // L2 already, so that load does not pay DRAM latency on the MMA's critical path
            int pc = c + STAGES, ptile = tile;
            if (pc >= a.chunks) { pc -= a.chunks; ptile += gridDim.x; }
            if (ptile < a.n_tiles) {
              const int nl = ptile == tile ? lvl : find_level(a, ptile);
              const TileCoord n = ptile == tile ? t : decode_tile(d, a.lv[nl], ptile - a.lv[nl].tile_begin, th, TW);
              const int ny = n.sy + dl * n.y0, nx = n.sx + dl * n.x0;
              const PairFrames nf = ptile == tile ? pf : pair_frames(d, n.b);
              tma_prefetch_4d(nf.alt ? &maps.x1_alt : &maps.x1[nl], pc * 64, nx, ny, nf.b1);
              tma_prefetch_4d(&maps.x2[nl], pc * 64, nx - r * dl, ny - r * dl, nf.b2);
            }
          }
          if (c == 0) STM_TRACE(k, 0);
          if (c == a.chunks - 1) STM_TRACE(k, 1);
          if (++s == STAGES) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)a.n_half);
      int s = 0;
      uint32_t phase = 0, tphase = 0;
      int k = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
        mbar_wait_relaxed(tmem_empty, tphase ^ 1u);      // epilogue has drained the accumulator of the previous tile
        tcgen05_fence_after();
        for (int c = 0; c < a.chunks; ++c) {
          mbar_wait_relaxed(&full_bar[s], phase);
          tcgen05_fence_after();
          if (c == 0) STM_TRACE(k, 2);
          if (c == a.chunks - 1) STM_TRACE(k, 3);
          const uint32_t base = smem_u32(smem + s * L.stage_bytes);
          const uint64_t adesc = umma_desc_sw128(base);
          const uint64_t bdesc0 = umma_desc_sw128(base + L.a_bytes);
          const uint64_t bdesc1 = umma_desc_sw128(base + L.a_bytes + a.n_half * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (c | k) != 0 ? 1u : 0u;
            umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc0 + (uint64_t)(2 * k), idesc, acc);
            umma_bf16(tmem_base + (uint32_t)a.n_half, adesc + (uint64_t)(2 * k), bdesc1 + (uint64_t)(2 * k), idesc, acc);
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full);
        tphase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 2 + DRAIN_WARPS) {
    // =============================== COPY warps (phase 0) ===============================
    // The concat copy of the two feature maps depends on nothing, so these warps stream it tile after tile
    // on their own: DRAM stays busy while the other warps wait for / drain the accumulator.
    const int ew = warp - 2 - DRAIN_WARPS;   // 0 .. COPY_WARPS-1
    const int et = tid - 64 - DRAIN_THREADS;
    constexpr int EPI_WARPS = COPY_WARPS, EPI_THREADS = COPY_WARPS * 32;
    constexpr bool relu = POST == 2;
    const bool copy_feats = (d.flags & STM_CORR_COPY_FEATS) != 0;
    const int fc = d.feat_c, foff = a.feat_off;
    int k = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles && copy_feats; tile += gridDim.x, ++k) {
      const CorrLevel& lv = a.lv[find_level(a, tile)];
      OT* out = reinterpret_cast<OT*>(lv.out);
      const bool nhwc = lv.out_stride_c == 1;
      const TileCoord t = decode_tile(d, lv, tile - lv.tile_begin, th, TW);
      const int64_t obase = t.b * lv.out_stride_n;
      const PairFrames pf = pair_frames(d, t.b);
      if (et == 0) STM_TRACE(k, 4);
      // ---- phase 0: concat copy of the two feature maps behind the correlation channels.  It does not
      //      depend on the accumulator, so it runs while the tensor core is still working on this tile. ----
      {
        if (a.fast) {
          // one warp per (map, tile row, half row): SEG consecutive pixels, lanes on consecutive 16-byte chunks, all
          // SEG loads of a lane in flight at once (the copy is DRAM-latency bound: ~5 KB in flight per warp)
          constexpr int SEG = TW / 2;
          const int cpr = fc >> 3;                       // 16-byte chunks per pixel and map
          const int n_items = 4 * th;
#pragma unroll 1
          for (int item = ew; item < n_items; item += EPI_WARPS) {
            const int m = item & 1, seg = (item >> 1) & 1, ty = item >> 2;
            const int y = t.sy + dl * (t.y0 + ty);
            const int x = t.sx + dl * (t.x0 + seg * SEG);                       // first pixel of the segment
            int nv = y < lv.h ? (lv.w - x + dl - 1) / dl : 0;                   // valid pixels in it
            nv = nv < 0 ? 0 : (nv > SEG ? SEG : nv);
            const bool alt = m == 0 && pf.alt;
            const int64_t fsw_ = m ? d.feat_b_stride_w : (alt ? d.feat_a_alt_stride_w : d.feat_a_stride_w);
            const int64_t fsw = fsw_ * dl;
            const __nv_bfloat16* src =
                m ? reinterpret_cast<const __nv_bfloat16*>(a.fb) + pf.b2 * d.feat_b_stride_n + y * d.feat_b_stride_h + x * d.feat_b_stride_w
                  : (alt ? reinterpret_cast<const __nv_bfloat16*>(d.feat_a_alt) + pf.b1 * d.feat_a_alt_stride_n + y * d.feat_a_alt_stride_h + x * d.feat_a_alt_stride_w
                         : reinterpret_cast<const __nv_bfloat16*>(a.fa) + pf.b1 * d.feat_a_stride_n + y * d.feat_a_stride_h + x * d.feat_a_stride_w);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(lv.out) + obase + y * lv.out_stride_h + x * lv.out_stride_w + foff + m * fc;
            const int64_t osw = lv.out_stride_w * dl;
            for (int ch = lane; ch < cpr; ch += 32) {
              uint4 v[SEG];
#pragma unroll
              for (int i = 0; i < SEG; ++i)
                if (i < nv) v[i] = __ldg(reinterpret_cast<const uint4*>(src + i * fsw) + ch);
#pragma unroll
              for (int i = 0; i < SEG; ++i) {
                if (i >= nv) continue;
                uint4 w = v[i];
                if (relu) { w.x = relu_bf16x2(w.x); w.y = relu_bf16x2(w.y); w.z = relu_bf16x2(w.z); w.w = relu_bf16x2(w.w); }
                reinterpret_cast<uint4*>(dst + i * osw)[ch] = w;
              }
            }
          }
        } else {
          // any layout / dtype, element-wise.  Channels-last output: one warp per pixel, lanes across channels
          // (coalesced); planar output: one thread per pixel, threads of a pixel split the channels.
          const bool f32in = d.feat_dtype == STM_F32;
          auto feat = [&](const void* fp, int64_t idx) {
            const float f = f32in ? reinterpret_cast<const float*>(fp)[idx] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(fp)[idx]);
            return relu ? fmaxf(f, 0.f) : f;
          };
          const int p_begin = nhwc ? ew : (et & 127), p_step = nhwc ? EPI_WARPS : 128;
          const int c_begin = nhwc ? lane : (et >> 7), c_step = nhwc ? 32 : EPI_THREADS / 128;
#pragma unroll 1
          for (int pix = p_begin; pix < npix; pix += p_step) {
            const int ty = pix / TW, tx = pix - ty * TW;
            const int y = t.sy + dl * (t.y0 + ty), x = t.sx + dl * (t.x0 + tx);
            if (y >= lv.h || x >= lv.w) continue;
            OT* op = out + obase + y * lv.out_stride_h + x * lv.out_stride_w;
            const void* fa_ = pf.alt ? d.feat_a_alt : a.fa;
            const int64_t ia = pf.alt ? pf.b1 * d.feat_a_alt_stride_n + y * d.feat_a_alt_stride_h + x * d.feat_a_alt_stride_w
                                      : pf.b1 * d.feat_a_stride_n + y * d.feat_a_stride_h + x * d.feat_a_stride_w;
            const int64_t ib = pf.b2 * d.feat_b_stride_n + y * d.feat_b_stride_h + x * d.feat_b_stride_w;
#pragma unroll 4
            for (int c = c_begin; c < fc; c += c_step) {
              op[(int64_t)(foff + c) * lv.out_stride_c] = from_f32<OT>(feat(fa_, ia + c));
              op[(int64_t)(foff + fc + c) * lv.out_stride_c] = from_f32<OT>(feat(a.fb, ib + c));
            }
            for (int c = PP + c_begin; c < foff; c += c_step) op[(int64_t)c * lv.out_stride_c] = from_f32<OT>(0.f);
          }
        }
      }

      if (et == 0) STM_TRACE(k, 5);
    }
  } else {
    // =============================== DRAIN warps (phases 1, 2) ===============================
    constexpr int EPI_WARPS = DRAIN_WARPS, EPI_THREADS = DRAIN_THREADS;
    const int et = tid - 64;                 // 0 .. EPI_THREADS-1
    const int ew = warp - 2;                 // 0 .. EPI_WARPS-1
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int hsel = ew >> 2;                // the four warps of a quarter interleave the region rows
    const int i_pix = q * 32 + lane;         // x1 pixel of this thread's TMEM lane: (py, px) in the patch
    const int py = i_pix / TW, px = i_pix - py * TW;
    const bool pix_live = i_pix < npix;
    const int py_lo = (q * 32) / TW;
    const int py_hi = min(th - 1, (q * 32 + 31) / TW);
    const int S = a.stage_stride;
    const float scale = d.scale, slope = d.leaky_slope;
    const int foff = a.feat_off;
    uint32_t tphase = 0;
    int k = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++k) {
      const CorrLevel& lv = a.lv[find_level(a, tile)];
      OT* out = reinterpret_cast<OT*>(lv.out);
      const bool nhwc = lv.out_stride_c == 1;
      const TileCoord t = decode_tile(d, lv, tile - lv.tile_begin, th, TW);
      const int64_t obase = t.b * lv.out_stride_n;
      mbar_wait(tmem_full, tphase);
      tcgen05_fence_after();
      if (et == 0) STM_TRACE(k, 6);
      // ---- phase 1: TMEM -> scale / leaky-ReLU / ReLU -> band -> staging[pixel][ph*P + pw] (output dtype) ----
      for (int R = py_lo + hsel; R <= py_hi + P - 1; R += EPI_WARPS / 4) {      // region rows this lane quarter needs
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(R * a.rw), v);
        tmem_ld_wait();
        const int ph = R - py;
        if (pix_live && ph >= 0 && ph < P) {
          OT* dst = staging + i_pix * S + ph * P - px;
#pragma unroll
          for (int j = 0; j < MAX_RW; ++j) {
            float f = __uint_as_float(v[j]) * scale;
            if (POST == 1) f = f > 0.f ? f : f * slope;
            if (POST == 2) f = fmaxf(f, 0.f);
            if ((unsigned)(j - px) < (unsigned)P) dst[j] = from_f32<OT>(f);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);          // the MMA warp may start the next tile
      tphase ^= 1u;
      named_barrier_sync(1, EPI_THREADS);              // staging complete
      if (et == 0) STM_TRACE(k, 7);

      // ---- phase 2: staging -> global, in the caller's layout ----
      if (a.fast) {
        // channels [0, feat_off) of every pixel: one half-warp per pixel, 16-byte chunks
        const int cpp = foff >> 3;
        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(lv.out) + obase;
        const int l16 = et & 15;
#pragma unroll 2
        for (int pix = et >> 4; pix < npix; pix += EPI_THREADS / 16) {
          const int ty = pix / TW, tx = pix - ty * TW;
          const int y = t.sy + dl * (t.y0 + ty), x = t.sx + dl * (t.x0 + tx);
          if (y < lv.h && x < lv.w) {
            const uint2* sp = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(staging) + pix * S);
            uint4* op = reinterpret_cast<uint4*>(ob + y * lv.out_stride_h + x * lv.out_stride_w);
            for (int ch = l16; ch < cpp; ch += 16) {
              const uint2 lo = sp[2 * ch], hi = sp[2 * ch + 1];
              op[ch] = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
          }
        }
      } else if (nhwc) {
        // channels-last: one warp per pixel; its P*P output channels are contiguous
#pragma unroll 1
        for (int pix = ew; pix < npix; pix += EPI_WARPS) {
          const int ty = pix / TW, tx = pix - ty * TW;
          const int y = t.sy + dl * (t.y0 + ty), x = t.sx + dl * (t.x0 + tx);
          if (y < lv.h && x < lv.w) {
            OT* orow = out + obase + y * lv.out_stride_h + x * lv.out_stride_w;
            const OT* srow = staging + pix * S;
            for (int k = lane; k < PP; k += 32) orow[k] = srow[k];
          }
        }
      } else {
        // planar (NCHW-like): every thread owns ONE patch pixel and walks the displacement planes; TW
        // consecutive pixels of a patch row are contiguous when out_stride_w == 1
        const int pix = et & 127;
        const int ty = pix / TW, tx = pix - ty * TW;
        const int y = t.sy + dl * (t.y0 + ty), x = t.sx + dl * (t.x0 + tx);
        if (pix < npix && y < lv.h && x < lv.w) {
          OT* op = out + obase + y * lv.out_stride_h + x * lv.out_stride_w;
          const OT* srow = staging + pix * S;
#pragma unroll 8
          for (int k = et >> 7; k < PP; k += EPI_THREADS / 128) op[(int64_t)k * lv.out_stride_c] = srow[k];
        }
      }
      named_barrier_sync(2, EPI_THREADS);              // staging free for the next tile
      if (et == 0) STM_TRACE(k, 8);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) STM_TRACE(4, 10);      // slot 58: teardown
#undef STM_TRACE
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int encode_nhwc_map(CUtensorMap* m, const void* ptr, int c, int w, int h, int b, int64_t sw, int64_t sh, int64_t sn, int box_w,
                    int box_h, int dl) {
  PFN_stm_encodeTiled enc = get_tensormap_encoder();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled unavailable"); return STM_ERR_CUDA; }
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)b};
  const cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)(box_w * dl), (cuuint32_t)(box_h * dl), 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)dl, (cuuint32_t)dl, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return STM_ERR_CUDA; }
  return STM_OK;
}

SmemAttrCache g_corr_smem_attr[12];    // one per instantiation (out dtype x tile width x post-op)

template <typename OT, int TW, int POST>
int launch_t(const CorrTcArgs& args, const CorrMaps& maps, int grid, int smem_bytes, cudaStream_t stream) {
  constexpr int slot = (sizeof(OT) == 4 ? 6 : 0) + (TW == 20 ? 3 : 0) + POST;
  const int rc = ensure_dynamic_smem(corr_tc_kernel<OT, TW, POST>, smem_bytes, g_corr_smem_attr[slot]);
  if (rc != STM_OK) return rc;
  corr_tc_kernel<OT, TW, POST><<<grid, NUM_THREADS, smem_bytes, stream>>>(args, maps);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

// tiles of one image (all dilation sub-lattices) for a th x tw patch
int tiles_per_image(const StmCorrDesc& d, int th, int tw) {
  const int dl = d.dilation_patch;
  int tiles = 0;
  for (int sy = 0; sy < dl; ++sy)
    for (int sx = 0; sx < dl; ++sx) {
      const int hs = (d.h - sy + dl - 1) / dl, ws = (d.w - sx + dl - 1) / dl;
      tiles += ((hs + th - 1) / th) * ((ws + tw - 1) / tw);
    }
  return tiles;
}

}  // namespace

bool corr_tc_supported(const StmCorrDesc& d, const char** why) {
  *why = "";
  if (d.dtype != STM_BF16) { *why = "dtype is not bf16"; return false; }
  if (d.c % 64 != 0 || d.c > 2048) { *why = "C not a multiple of 64 (or > 2048)"; return false; }
  if (d.patch > MAX_P || (d.patch & 1) == 0) { *why = "patch_size > 11"; return false; }
  if (d.dilation_patch * (20 + d.patch - 1) > 256) { *why = "dilation_patch too large for a TMA box"; return false; }
  if ((d.x1_stride_n | d.x1_stride_h | d.x1_stride_w | d.x2_stride_n | d.x2_stride_h | d.x2_stride_w) & 7) {
    *why = "x1 / x2 strides not multiples of 8 elements";
    return false;
  }
  if (d.x1_stride_w <= 0 || d.x1_stride_h <= 0 || d.x2_stride_w <= 0 || d.x2_stride_h <= 0 ||
      (d.batch > 1 && (d.x1_stride_n <= 0 || d.x2_stride_n <= 0))) { *why = "non-positive strides"; return false; }
  if (d.batch < 1) { *why = "empty batch"; return false; }
  if (get_tensormap_encoder() == nullptr) { *why = "cuTensorMapEncodeTiled unavailable"; return false; }
  return true;
}

// One launch over n feature maps (levels) that share C, patch, dilation, dtypes and post-ops: descs[i] carries level
// i's size, strides and (for i == 0 only) the concat / pair-indexing fields.
int launch_corr_tc_multi(const StmCorrDesc* descs, const void* const* x1s, const void* const* x2s, const void* fa, const void* fb,
                         void* const* outs, int n, cudaStream_t stream) {
  const StmCorrDesc& d = descs[0];
  if (n < 1 || n > MAX_LEVELS) { set_error("1..%d feature maps per launch, got %d", MAX_LEVELS, n); return STM_ERR_INVALID_ARGUMENT; }
  for (int i = 0; i < n; ++i) {
    if (((uintptr_t)x1s[i] & 15) || ((uintptr_t)x2s[i] & 15)) {
      // descriptors need 16-byte aligned bases; rare (sliced tensors) -> CUDA-core kernel, which knows nothing
      // about pair indexing or grouped launches
      if (d.x1_index != nullptr || n > 1) {
        set_error("pair-indexed / grouped correlation needs 16-byte aligned x1 / x2 (tcgen05 backend only)");
        return STM_ERR_UNSUPPORTED;
      }
      return launch_corr_simt(d, x1s[0], x2s[0], fa, fb, outs[0], stream);
    }
  }
  CorrTcArgs args;
  memset(&args, 0, sizeof(args));
  args.d = d;
  args.fa = fa; args.fb = fb;
  args.trace = nullptr;
  args.l2_prefetch = 1;
#ifdef STM_DCN_EXPERIMENTS   // profiling builds only (tools/corr_trace.py); the product library has no environment knobs
  if (const char* e = getenv("STM_CORR_L2PF")) args.l2_prefetch = atoi(e);
  if (const char* e = getenv("STM_DEBUG_BUF")) args.trace = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
#endif
  const int P = d.patch, dl = d.dilation_patch;
  // tile shape: fewest tiles wins (6x20 tiles a 24x40 map exactly: 8 tiles instead of 9); ties -> 8x16
  int tw = 16, th = 8;
  {
    int t816 = 0, t620 = 0;
    for (int i = 0; i < n; ++i) { t816 += tiles_per_image(descs[i], 8, 16) * descs[i].batch; t620 += tiles_per_image(descs[i], 6, 20) * descs[i].batch; }
    if (t620 < t816) { tw = 20; th = 6; }
  }
  args.th = th;
  args.rh = th + P - 1;
  args.rw = tw + P - 1;
  const int n_region = args.rh * args.rw;
  args.n_half = (((n_region + 1) / 2) + 15) & ~15;
  args.chunks = d.c / 64;
  int need_cols = 2 * args.n_half;
  if ((args.rh - 1) * args.rw + 32 > need_cols) need_cols = (args.rh - 1) * args.rw + 32;
  int cols = 32;
  while (cols < need_cols) cols <<= 1;
  if (cols > 512) { set_error("tcgen05 correlation: region of %d pixels does not fit TMEM", n_region); return STM_ERR_UNSUPPORTED; }
  args.tmem_cols = cols;
  args.n_levels = n;
  args.n_tiles = 0;
  for (int i = 0; i < n; ++i) {
    const StmCorrDesc& q = descs[i];
    CorrLevel& lv = args.lv[i];
    lv.h = q.h; lv.w = q.w;
    lv.tile_begin = args.n_tiles;
    lv.tiles_per_image = tiles_per_image(q, th, tw);
    lv.out = outs[i];
    lv.out_stride_n = q.out_stride_n; lv.out_stride_c = q.out_stride_c; lv.out_stride_h = q.out_stride_h; lv.out_stride_w = q.out_stride_w;
    args.n_tiles += lv.tiles_per_image * q.batch;
  }
  if (args.n_tiles == 0) return STM_OK;
  const int oes = d.out_dtype == STM_F32 ? 4 : 2;
  const bool copy = (d.flags & STM_CORR_COPY_FEATS) != 0;
  // first channel behind the P*P correlation channels: the concat's feature block, or (without a concat) the end of a
  // zero-padded cost-volume row — feat_c_offset = 128 for P = 11 gives 256-byte channels-last rows
  args.feat_off = d.feat_c_offset > P * P ? d.feat_c_offset : P * P;
  // aligned channels-last epilogue: bf16 everywhere, every pixel row / feature block / feature row 16-byte aligned
  bool fast = d.out_dtype == STM_BF16 && (args.feat_off & 7) == 0 && (copy || args.feat_off > P * P);
  for (int i = 0; i < n && fast; ++i)
    fast = descs[i].out_stride_c == 1 && ((descs[i].out_stride_n | descs[i].out_stride_h | descs[i].out_stride_w) & 7) == 0 &&
           (((uintptr_t)outs[i]) & 15) == 0;
  if (copy)
    fast = fast && d.feat_dtype == STM_BF16 && (d.feat_c & 7) == 0 &&
           (((d.feat_a_stride_n | d.feat_a_stride_h | d.feat_a_stride_w | d.feat_b_stride_n | d.feat_b_stride_h | d.feat_b_stride_w) & 7) == 0) &&
           ((((uintptr_t)fa | (uintptr_t)fb) & 15) == 0) &&
           (!(d.x1_index != nullptr && d.alt_frames > 0) ||
            ((((uintptr_t)d.feat_a_alt) & 15) == 0 && ((d.feat_a_alt_stride_n | d.feat_a_alt_stride_h | d.feat_a_alt_stride_w) & 7) == 0));
  args.fast = fast ? 1 : 0;
  if (!fast && !copy && args.feat_off > P * P) {
    set_error("a zero-padded cost volume (feat_c_offset > patch^2 without COPY_FEATS) needs channels-last bf16 output with 16-byte aligned rows");
    return STM_ERR_UNSUPPORTED;
  }
  // staging row stride (elements).  fast: feat_off elements + 4 (rows stay 8-byte aligned for LDS.64, and the
  // 2-word skew spreads the band scatter over the banks); generic: an odd number of 32-bit words per pixel
  args.stage_stride = args.fast ? args.feat_off + 4 : (oes == 4 ? (P * P + 2) : ((P * P + 3) & ~1));
  const SmemPlan L(args.n_half, args.stage_stride, oes);
  if (L.total > 227 * 1024) { set_error("tcgen05 correlation: shared memory %d B over the limit", L.total); return STM_ERR_UNSUPPORTED; }

  // frame counts of the tensors behind the maps (pair indexing addresses frames, not pairs)
  const bool indexed = d.x1_index != nullptr;
  CorrMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < n; ++i) {
    const StmCorrDesc& q = descs[i];
    const int nb1 = indexed ? q.x1_frames : q.batch, nb2 = indexed ? q.x2_frames : q.batch;
    int rc = encode_nhwc_map(&maps.x1[i], x1s[i], q.c, q.w, q.h, nb1, q.x1_stride_w, q.x1_stride_h,
                             nb1 > 1 ? q.x1_stride_n : (int64_t)q.h * q.x1_stride_h, tw, th, dl);
    if (rc != STM_OK) return rc;
    rc = encode_nhwc_map(&maps.x2[i], x2s[i], q.c, q.w, q.h, nb2, q.x2_stride_w, q.x2_stride_h,
                         nb2 > 1 ? q.x2_stride_n : (int64_t)q.h * q.x2_stride_h, args.rw, args.rh, dl);
    if (rc != STM_OK) return rc;
  }
  maps.x1_alt = maps.x1[0];
  if (indexed && d.alt_frames > 0) {
    if (((uintptr_t)d.x1_alt & 15) || ((d.x1_alt_stride_n | d.x1_alt_stride_h | d.x1_alt_stride_w) & 7)) {
      set_error("x1_alt must be 16-byte aligned with strides that are multiples of 8 elements");
      return STM_ERR_INVALID_ARGUMENT;
    }
    const int rc = encode_nhwc_map(&maps.x1_alt, d.x1_alt, d.c, d.w, d.h, d.alt_frames, d.x1_alt_stride_w, d.x1_alt_stride_h,
                                   d.alt_frames > 1 ? d.x1_alt_stride_n : (int64_t)d.h * d.x1_alt_stride_h, tw, th, dl);
    if (rc != STM_OK) return rc;
  }
  const int grid = args.n_tiles < device_sm_count() ? args.n_tiles : device_sm_count();
  const int post = (d.flags & STM_CORR_RELU) ? 2 : ((d.flags & STM_CORR_LEAKY_RELU) ? 1 : 0);
#define STM_CORR_LAUNCH(OT_, TW_)                                                                     \
  do {                                                                                                \
    if (post == 2) return launch_t<OT_, TW_, 2>(args, maps, grid, L.total, stream);                        \
    if (post == 1) return launch_t<OT_, TW_, 1>(args, maps, grid, L.total, stream);                        \
    return launch_t<OT_, TW_, 0>(args, maps, grid, L.total, stream);                                       \
  } while (0)
  if (tw == 20) {
    if (d.out_dtype == STM_F32) STM_CORR_LAUNCH(float, 20);
    STM_CORR_LAUNCH(__nv_bfloat16, 20);
  }
  if (d.out_dtype == STM_F32) STM_CORR_LAUNCH(float, 16);
  STM_CORR_LAUNCH(__nv_bfloat16, 16);
#undef STM_CORR_LAUNCH
}

int launch_corr_tc(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                   cudaStream_t stream) {
  return launch_corr_tc_multi(&d, &x1, &x2, fa, fb, &out, 1, stream);
}

}  // namespace stm
