// Correlation cost volume on tcgen05 tensor cores.
//
// out[b, ph*P+pw, y, x] = post(scale * <x1[b,y,x,:], x2[b, y+(ph-r)d, x+(pw-r)d, :]>)
//
// The C-contraction is a GEMM between a PATCH of x1 pixels and the HALO REGION of x2 pixels around it:
//   tile    = 8 x 16 x1 pixels                (M = 128 rows,       K-major [pixel][channel])
//   region  = (8+P-1) x (16+P-1) x2 pixels    (N = 468 for P = 11, K-major [pixel][channel])
//   D[128, N] = X1_tile * X2_region^T, fp32 in TMEM; the P*P wanted displacements of pixel (py, px) are
//   the columns (py+ph) * RW + (px+pw)  — a lane-dependent band that the epilogue picks out.
// Both operands arrive by 4-D TMA straight from the NHWC tensors (box {64 ch, w, h, 1}, 128B swizzle);
// out-of-map halo pixels are zero-filled by TMA, which IS the operator's border rule.  A dilated patch
// (dilation_patch = d) is d*d independent undilated problems on the sub-lattices (y % d, x % d); TMA
// element strides {1, d, d, 1} load a sub-lattice directly, so the same kernel serves every d.
//
// Per persistent CTA (one per SM):  warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner),
// warps 2-17 epilogue: tcgen05.ld -> scale / leaky-ReLU / ReLU -> band scatter into a padded smem
// staging tile (conflict-free) -> coalesced stores in the caller's layout (NCHW or NHWC, any strides),
// then the optional concat copy of the two T2S feature maps (TF_utils.py:30-31).
//
// Replaces correlation_cuda_forward_kernel (+ its two NHWC permute copies) of
// spatial_correlation_sampler and the elementwise tail of correlate()/CandidateShift
// (reference track_to_segment_head.py:53-62, TF_utils.py:28-31).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace stm {
namespace {

using namespace tc;

constexpr int PATCH_H = 8, PATCH_W = 16;      // x1 pixels per tile (M = 128)
constexpr int STAGES = 2;
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + EPI_THREADS;
constexpr int MAX_RW = 26;                    // region width for P = 11

struct CorrTcArgs {
  StmCorrDesc d;
  const void* fa;
  const void* fb;
  void* out;
  int32_t n_tiles, tiles_per_image;
  int32_t rh, rw;          // region height / width in (sub-lattice) pixels
  int32_t n_half;          // N of each of the two MMAs (multiple of 16, <= 256)
  int32_t chunks;          // C / 64
  int32_t tmem_cols;
  int32_t stage_stride;    // elements per pixel in the staging tile
  int32_t slot_bytes;      // per-warp row slot for the vectorised feature copy (0 = scalar copy)
  long long* dbg;          // optional clock64 trace (STM_DEBUG_BUF), 32 slots per CTA
};

struct SmemPlan {
  int a_bytes, b_bytes, stage_bytes, staging, slots, bars, total;
  __host__ __device__ SmemPlan(int n_half, int stage_stride, int out_esize, int slot_bytes) {
    a_bytes = PATCH_H * PATCH_W * 128;
    b_bytes = 2 * n_half * 128;
    stage_bytes = a_bytes + b_bytes;
    staging = STAGES * stage_bytes;
    slots = staging + ((PATCH_H * PATCH_W * stage_stride * out_esize + 15) & ~15);
    bars = slots + EPI_WARPS * slot_bytes;
    total = bars + 128 + 1024;
  }
};

struct TileCoord {
  int b, sy, sx, y0, x0;   // batch, sub-lattice phase, first pixel of the patch IN SUB-LATTICE coordinates
};

__device__ __forceinline__ TileCoord decode_tile(const StmCorrDesc& d, int tile, int tiles_per_image) {
  TileCoord t;
  t.b = tile / tiles_per_image;
  int rem = tile - t.b * tiles_per_image;
  const int dl = d.dilation_patch;
  t.sy = t.sx = t.y0 = t.x0 = 0;
  for (int sy = 0; sy < dl; ++sy)
    for (int sx = 0; sx < dl; ++sx) {
      const int hs = (d.h - sy + dl - 1) / dl, ws = (d.w - sx + dl - 1) / dl;
      const int ty = (hs + PATCH_H - 1) / PATCH_H, tx = (ws + PATCH_W - 1) / PATCH_W;
      const int n = ty * tx;
      if (rem >= 0 && rem < n) {
        t.sy = sy; t.sx = sx;
        t.y0 = (rem / tx) * PATCH_H;
        t.x0 = (rem % tx) * PATCH_W;
        rem = -1;
      } else if (rem >= 0) {
        rem -= n;
      }
    }
  return t;
}

template <typename OT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_tc_kernel(const __grid_constant__ CorrTcArgs a, const __grid_constant__ CUtensorMap tmap_x1,
               const __grid_constant__ CUtensorMap tmap_x2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemPlan L(a.n_half, a.stage_stride, (int)sizeof(OT), a.slot_bytes);
  OT* staging = reinterpret_cast<OT*>(smem + L.staging);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const StmCorrDesc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = d.patch, PP = P * P, r = P / 2, dl = d.dilation_patch;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_x1);
    prefetch_tensormap(&tmap_x2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 32 : nullptr;
#define STM_DBG(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  if (tid == 0) STM_DBG(0);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)((PATCH_H * PATCH_W + a.rh * a.rw) * 128);
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(d, tile, a.tiles_per_image);
        const int fy = t.sy + dl * t.y0, fx = t.sx + dl * t.x0;           // full-resolution coords of the patch origin
        for (int c = 0; c < a.chunks; ++c) {
          mbar_wait_relaxed(&empty_bar[s], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          uint8_t* st = smem + s * L.stage_bytes;
          tma_load_4d(st, &tmap_x1, &full_bar[s], c * 64, fx, fy, t.b);
          tma_load_4d(st + L.a_bytes, &tmap_x2, &full_bar[s], c * 64, fx - r * dl, fy - r * dl, t.b);
          if (tile == (int)blockIdx.x && c < 4) STM_DBG(1 + c);
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
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait_relaxed(tmem_empty, tphase ^ 1u);      // epilogue has drained the accumulator of the previous tile
        tcgen05_fence_after();
        for (int c = 0; c < a.chunks; ++c) {
          mbar_wait_relaxed(&full_bar[s], phase);
          tcgen05_fence_after();
          if (tile == (int)blockIdx.x && c < 4) STM_DBG(5 + c);
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
        if (tile == (int)blockIdx.x) STM_DBG(9);
        tphase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE ===============================
    const int et = tid - 64;                 // 0 .. EPI_THREADS-1
    const int ew = warp - 2;                 // 0 .. EPI_WARPS-1
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int hsel = ew >> 2;                // the four warps of a quarter interleave the region rows
    const int i_pix = q * 32 + lane;         // x1 pixel of this thread's TMEM lane: (py, px) in the patch
    const int px = i_pix & (PATCH_W - 1);
    const int pyl = lane >> 4;               // py - 2q
    const int S = a.stage_stride;
    const bool leaky = (d.flags & STM_CORR_LEAKY_RELU) != 0, relu = (d.flags & STM_CORR_RELU) != 0;
    const bool copy_feats = (d.flags & STM_CORR_COPY_FEATS) != 0;
    const int fc = d.feat_c;
    OT* out = reinterpret_cast<OT*>(a.out);
    const bool nhwc = d.out_stride_c == 1;
    const bool vec_feats = copy_feats && a.slot_bytes > 0;     // host checked layout / alignment / dtype
    uint8_t* slot = smem + L.slots + ew * a.slot_bytes;
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(d, tile, a.tiles_per_image);
      const int64_t obase = t.b * d.out_stride_n;

      // ---- phase 0: concat copy of the two feature maps behind the P*P correlation channels.  It does not
      //      depend on the accumulator, so it runs while the tensor core is still working on this tile. ----
      if (copy_feats) {
        if (vec_feats) {
          // One warp per pixel.  The 2*fc bf16 of a pixel form ONE contiguous run in `out`, starting at a
          // 2-byte-aligned address.  They are assembled in a per-warp smem slot with the SAME 16-byte phase
          // as the destination, so the bulk leaves as aligned 16-byte stores (head / tail as 2-byte stores).
          constexpr int PIX_PER_ITER = 4;
#pragma unroll 1
          for (int p0 = ew; p0 < PATCH_H * PATCH_W; p0 += PIX_PER_ITER * EPI_WARPS) {
            uint4 fv[PIX_PER_ITER][2];
            int64_t ob[PIX_PER_ITER];
            bool ok[PIX_PER_ITER];
#pragma unroll
            for (int u = 0; u < PIX_PER_ITER; ++u) {
              const int pix = p0 + u * EPI_WARPS;
              const int y = t.sy + dl * (t.y0 + (pix >> 4)), x = t.sx + dl * (t.x0 + (pix & 15));
              ok[u] = pix < PATCH_H * PATCH_W && y < d.h && x < d.w;
              ob[u] = obase + y * d.out_stride_h + x * d.out_stride_w + PP;       // first feature element
              const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(a.fa) + t.b * d.feat_a_stride_n + y * d.feat_a_stride_h + x * d.feat_a_stride_w;
              const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(a.fb) + t.b * d.feat_b_stride_n + y * d.feat_b_stride_h + x * d.feat_b_stride_w;
              const int c0 = lane * 8;                                // fc <= 256: one 16-byte vector per lane and map
              const bool ld = ok[u] && c0 < fc;
              fv[u][0] = ld ? __ldg(reinterpret_cast<const uint4*>(pa + c0)) : make_uint4(0u, 0u, 0u, 0u);
              fv[u][1] = ld ? __ldg(reinterpret_cast<const uint4*>(pb + c0)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < PIX_PER_ITER; ++u) {
              if (!ok[u]) continue;                                  // warp-uniform
              uint8_t* gdst = reinterpret_cast<uint8_t*>(out) + ob[u] * 2;
              const int ph16 = (int)((uintptr_t)gdst & 15);          // destination phase inside a 16-byte line
              uint16_t* s16 = reinterpret_cast<uint16_t*>(slot + ph16);
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                  const int c0 = lane * 8;
                  if (c0 >= fc) continue;
                  const uint4 vv = fv[u][m];
                  uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
                  if (relu) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {        // bf16 ReLU: clear the halves whose sign bit is set
                      const uint32_t neg = w[i] & 0x80008000u;
                      w[i] &= ~(((neg >> 15) & 0x00010001u) * 0xffffu);
                    }
                  }
                  uint16_t* sd = s16 + m * fc + c0;
                  if ((ph16 & 2) == 0) {                 // 4-byte aligned in the slot
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(sd);
                    d32[0] = w[0]; d32[1] = w[1]; d32[2] = w[2]; d32[3] = w[3];
                  } else {
                    sd[0] = (uint16_t)(w[0] & 0xffffu);
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(sd + 1);
                    d32[0] = __funnelshift_r(w[0], w[1], 16);
                    d32[1] = __funnelshift_r(w[1], w[2], 16);
                    d32[2] = __funnelshift_r(w[2], w[3], 16);
                    sd[7] = (uint16_t)(w[3] >> 16);
                  }
                }
              __syncwarp();
              const int nbytes = 4 * fc;                             // 2 maps * fc * 2 B
              const int head = (16 - ph16) & 15;                     // bytes up to the first 16-byte boundary
              const int body = (nbytes - head) >> 4;                 // full 16-byte chunks
              const int tail0 = head + (body << 4);
              if (2 * lane < head)
                *reinterpret_cast<uint16_t*>(gdst + 2 * lane) = *reinterpret_cast<const uint16_t*>(slot + ph16 + 2 * lane);
              for (int ch = lane; ch < body; ch += 32)
                *reinterpret_cast<uint4*>(gdst + head + (ch << 4)) = *reinterpret_cast<const uint4*>(slot + ph16 + head + (ch << 4));
              if (tail0 + 2 * lane < nbytes)
                *reinterpret_cast<uint16_t*>(gdst + tail0 + 2 * lane) = *reinterpret_cast<const uint16_t*>(slot + ph16 + tail0 + 2 * lane);
              __syncwarp();
            }
          }
        } else {
          // any layout / dtype: every thread owns one patch pixel (threads of a pixel split the channels)
          const int pix = et & 127;
          const int y = t.sy + dl * (t.y0 + (pix >> 4)), x = t.sx + dl * (t.x0 + (pix & 15));
          if (y < d.h && x < d.w) {
            OT* op = out + obase + y * d.out_stride_h + x * d.out_stride_w;
            const bool f32in = d.feat_dtype == STM_F32;
            const int64_t ia = t.b * d.feat_a_stride_n + y * d.feat_a_stride_h + x * d.feat_a_stride_w;
            const int64_t ib = t.b * d.feat_b_stride_n + y * d.feat_b_stride_h + x * d.feat_b_stride_w;
#pragma unroll 8
            for (int c = et >> 7; c < 2 * fc; c += EPI_THREADS / 128) {
              const bool second = c >= fc;
              const int64_t src = (second ? ib : ia) + (second ? c - fc : c);
              const void* fp = second ? a.fb : a.fa;
              float f = f32in ? reinterpret_cast<const float*>(fp)[src] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(fp)[src]);
              if (relu) f = fmaxf(f, 0.f);
              op[(int64_t)(PP + c) * d.out_stride_c] = from_f32<OT>(f);
            }
          }
        }
      }

      mbar_wait(tmem_full, tphase);
      tcgen05_fence_after();
      if (et == 0 && tile == (int)blockIdx.x) STM_DBG(10);
      // ---- phase 1: TMEM -> scale / leaky-ReLU / ReLU -> band -> staging[pixel][ph*P + pw] (output dtype) ----
      for (int rr = hsel; rr <= P; rr += EPI_WARPS / 4) {          // region rows 2q + rr, rr in [0, P]
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((2 * q + rr) * a.rw), v);
        tmem_ld_wait();
        const int ph = rr - pyl;
        if (ph >= 0 && ph < P) {
          OT* dst = staging + i_pix * S + ph * P - px;
#pragma unroll
          for (int j = 0; j < MAX_RW; ++j) {
            float f = __uint_as_float(v[j]) * d.scale;
            if (leaky) f = f > 0.f ? f : f * d.leaky_slope;
            if (relu) f = fmaxf(f, 0.f);
            if ((unsigned)(j - px) < (unsigned)P) dst[j] = from_f32<OT>(f);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);          // the MMA warp may start the next tile
      tphase ^= 1u;
      named_barrier_sync(1, EPI_THREADS);              // staging complete
      if (et == 0 && tile == (int)blockIdx.x) STM_DBG(11);

      // ---- phase 2: staging -> global, in the caller's layout ----
      if (nhwc) {
        // channels-last: one warp per pixel; its P*P output channels are contiguous
#pragma unroll 1
        for (int p0 = ew; p0 < PATCH_H * PATCH_W; p0 += 4 * EPI_WARPS) {
          OT vals[4][4];
          OT* orow[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int pix = p0 + u * EPI_WARPS;
            const int y = t.sy + dl * (t.y0 + (pix >> 4)), x = t.sx + dl * (t.x0 + (pix & 15));
            ok[u] = pix < PATCH_H * PATCH_W && y < d.h && x < d.w;
            orow[u] = out + obase + y * d.out_stride_h + x * d.out_stride_w;
            const OT* srow = staging + (pix & (PATCH_H * PATCH_W - 1)) * S;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = lane + 32 * i;
              if (k < PP) vals[u][i] = srow[k];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = lane + 32 * i;
              if (ok[u] && k < PP) orow[u][k] = vals[u][i];
            }
        }
      } else {
        // planar (NCHW-like): every thread owns ONE patch pixel and walks the displacement planes; 16
        // consecutive pixels of a patch row are contiguous when out_stride_w == 1
        const int pix = et & 127;
        const int y = t.sy + dl * (t.y0 + (pix >> 4)), x = t.sx + dl * (t.x0 + (pix & 15));
        if (y < d.h && x < d.w) {
          OT* op = out + obase + y * d.out_stride_h + x * d.out_stride_w;
          const OT* srow = staging + pix * S;
#pragma unroll 8
          for (int k = et >> 7; k < PP; k += EPI_THREADS / 128) op[(int64_t)k * d.out_stride_c] = srow[k];
        }
      }
      named_barrier_sync(2, EPI_THREADS);              // staging free for the next tile
      if (et == 0 && tile == (int)blockIdx.x) STM_DBG(12);
      if (et == 0 && tile == (int)blockIdx.x + (int)gridDim.x) STM_DBG(13);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) STM_DBG(14);
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
#undef STM_DBG
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

int sm_count_corr() {
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

template <typename OT>
int launch_t(const CorrTcArgs& args, const CUtensorMap& m1, const CUtensorMap& m2, int grid, int smem_bytes, cudaStream_t stream) {
  static int configured = 0;
  if (configured < smem_bytes) {
    STM_CUDA_OK(cudaFuncSetAttribute(corr_tc_kernel<OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = smem_bytes;
  }
  corr_tc_kernel<OT><<<grid, NUM_THREADS, smem_bytes, stream>>>(args, m1, m2);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace

bool corr_tc_supported(const StmCorrDesc& d, const char** why) {
  *why = "";
  if (d.dtype != STM_BF16) { *why = "dtype is not bf16"; return false; }
  if (d.c % 64 != 0 || d.c > 2048) { *why = "C not a multiple of 64 (or > 2048)"; return false; }
  if (d.patch > 11 || (d.patch & 1) == 0) { *why = "patch_size > 11"; return false; }
  if (d.dilation_patch * (PATCH_W + d.patch - 1) > 256) { *why = "dilation_patch too large for a TMA box"; return false; }
  if ((d.x1_stride_n | d.x1_stride_h | d.x1_stride_w | d.x2_stride_n | d.x2_stride_h | d.x2_stride_w) & 7) {
    *why = "x1 / x2 strides not multiples of 8 elements";
    return false;
  }
  if (d.x1_stride_w <= 0 || d.x1_stride_h <= 0 || d.x2_stride_w <= 0 || d.x2_stride_h <= 0 ||
      (d.batch > 1 && (d.x1_stride_n <= 0 || d.x2_stride_n <= 0))) { *why = "non-positive strides"; return false; }
  if (d.batch < 1) { *why = "empty batch"; return false; }
  if (get_tensormap_encoder() == nullptr) { *why = "cuTensorMapEncodeTiled unavailable"; return false; }
  if (const char* e = getenv("STM_CORR_FORCE_SIMT")) {
    if (atoi(e) != 0) { *why = "STM_CORR_FORCE_SIMT"; return false; }
  }
  return true;
}

int launch_corr_tc(const StmCorrDesc& d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                   cudaStream_t stream) {
  if (((uintptr_t)x1 & 15) || ((uintptr_t)x2 & 15)) {
    // descriptors need 16-byte aligned bases; rare (sliced tensors) -> CUDA-core kernel
    return launch_corr_simt(d, x1, x2, fa, fb, out, stream);
  }
  CorrTcArgs args;
  args.d = d;
  args.fa = fa; args.fb = fb; args.out = out;
  args.dbg = nullptr;
  if (const char* e = getenv("STM_DEBUG_BUF")) args.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  const int P = d.patch, dl = d.dilation_patch;
  args.rh = PATCH_H + P - 1;
  args.rw = PATCH_W + P - 1;
  const int n_region = args.rh * args.rw;
  args.n_half = (((n_region + 1) / 2) + 15) & ~15;
  args.chunks = d.c / 64;
  int need_cols = 2 * args.n_half;
  if ((args.rh - 1) * args.rw + 32 > need_cols) need_cols = (args.rh - 1) * args.rw + 32;
  int cols = 32;
  while (cols < need_cols) cols <<= 1;
  args.tmem_cols = cols;
  int tiles = 0;
  for (int sy = 0; sy < dl; ++sy)
    for (int sx = 0; sx < dl; ++sx) {
      const int hs = (d.h - sy + dl - 1) / dl, ws = (d.w - sx + dl - 1) / dl;
      tiles += ((hs + PATCH_H - 1) / PATCH_H) * ((ws + PATCH_W - 1) / PATCH_W);
    }
  args.tiles_per_image = tiles;
  args.n_tiles = tiles * d.batch;
  if (args.n_tiles == 0) return STM_OK;
  const int oes = d.out_dtype == STM_F32 ? 4 : 2;
  // staging row stride (elements): odd number of 32-bit words per pixel keeps the band scatter spread over the banks
  args.stage_stride = oes == 4 ? (P * P + 2) : ((P * P + 3) & ~1);
  // vectorised concat copy: bf16 everywhere, channels-last output, 16-byte aligned feature rows
  args.slot_bytes = 0;
  if ((d.flags & STM_CORR_COPY_FEATS) && d.out_stride_c == 1 && d.out_dtype == STM_BF16 && d.feat_dtype == STM_BF16 &&
      (d.feat_c & 7) == 0 && d.feat_c <= 256 &&
      (((d.feat_a_stride_n | d.feat_a_stride_h | d.feat_a_stride_w | d.feat_b_stride_n | d.feat_b_stride_h | d.feat_b_stride_w) & 7) == 0) &&
      ((((uintptr_t)fa | (uintptr_t)fb) & 15) == 0) && (((uintptr_t)out & 1) == 0))
    args.slot_bytes = ((4 * d.feat_c + 16 + 15) & ~15) + 16;
  SmemPlan L(args.n_half, args.stage_stride, oes, args.slot_bytes);
  if (L.total > 227 * 1024 && args.slot_bytes > 0) {     // no room for the row slots: scalar concat copy
    args.slot_bytes = 0;
    L = SmemPlan(args.n_half, args.stage_stride, oes, 0);
  }
  if (L.total > 227 * 1024) { set_error("tcgen05 correlation: shared memory %d B over the limit", L.total); return STM_ERR_UNSUPPORTED; }

  CUtensorMap m1, m2;
  int rc = encode_nhwc_map(&m1, x1, d.c, d.w, d.h, d.batch, d.x1_stride_w, d.x1_stride_h, d.batch > 1 ? d.x1_stride_n : (int64_t)d.h * d.x1_stride_h,
                           PATCH_W, PATCH_H, dl);
  if (rc != STM_OK) return rc;
  rc = encode_nhwc_map(&m2, x2, d.c, d.w, d.h, d.batch, d.x2_stride_w, d.x2_stride_h, d.batch > 1 ? d.x2_stride_n : (int64_t)d.h * d.x2_stride_h,
                       args.rw, args.rh, dl);
  if (rc != STM_OK) return rc;
  const int grid = args.n_tiles < sm_count_corr() ? args.n_tiles : sm_count_corr();
  if (d.out_dtype == STM_F32) return launch_t<float>(args, m1, m2, grid, L.total, stream);
  return launch_t<__nv_bfloat16>(args, m1, m2, grid, L.total, stream);
}

}  // namespace stm
