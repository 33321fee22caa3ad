// Deformable convolution forward on CUDA cores: any shape, fp32 math.
//
// This is the general path behind stm_deform_conv2d_fwd (groups, odd channel counts,
// fp32 storage — the <=1e-4 parity path).  The bf16 hot path for the STMask shapes is the
// tcgen05 kernel in dcn_tc.cu.
//
// Implicit GEMM  Y[M, Cout] = A[M, K] * W[Cout, K]^T,  M = B*Ho*Wo, K = kh*kw*Cin/groups.
// A is never materialised: for every (tap, deformable group) the block computes the four
// bilinear corner weights/offsets of its 64 output pixels once, then gathers 16-channel
// slices of A into shared memory (NHWC => a corner is a contiguous channel vector) and
// multiplies them with the matching OHWI weight slice.  Replaces
// modulated_deformable_im2col + SGEMM of dcn_v2 / mmcv (reference backbone.py:45,
// Featurealign.py:72).
#include "common.cuh"

namespace stm {
namespace {

constexpr int BM = 64;   // output pixels per block
constexpr int BN = 64;   // output channels per block
constexpr int KC = 16;   // input channels per smem slice
constexpr int NT = 256;  // threads

template <typename T> struct Vec4;  // 4 consecutive output channels
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<__nv_bfloat16> { using type = uint2; };

template <typename T, typename OT>
__global__ void __launch_bounds__(NT) dcn_simt_kernel(const __grid_constant__ DcnParams p) {
  __shared__ __align__(16) float As[KC][BM];
  __shared__ __align__(16) float Ws[KC][BN + 4];
  __shared__ float4 meta_w[BM];
  __shared__ int4 meta_o[BM];
  __shared__ int64_t pix_xbase[BM];   // element offset of x[b, 0, 0, 0]
  __shared__ int64_t pix_ybase[BM];   // element offset of y[b, ho, wo, 0]; -1 => row out of range
  __shared__ int pix_b[BM], pix_ho[BM], pix_wo[BM];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  // which problem does this M tile belong to?
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < p.n_probs; ++i)
    if (tile >= p.prob[i].tile_begin) pi = i;
  const DcnProblemDev& pr = p.prob[pi];
  const int m0 = (tile - pr.tile_begin) * BM;

  const int grp = blockIdx.z;
  const int cpg = p.in_c / p.groups;           // input channels per weight group
  const int opg = p.out_c / p.groups;
  const int cpd = p.in_c / p.dg;               // input channels per deformable group
  const int gc0 = grp * cpg, gc1 = gc0 + cpg;  // this group's input channel range
  const int oc0 = grp * opg + blockIdx.y * BN; // first output channel of the tile
  const int oc_end = grp * opg + opg;
  const int K = p.kh * p.kw;

  if (tid < BM) {
    const int m = m0 + tid;
    int b = 0, ho = 0, wo = 0;
    int64_t yb = -1;
    if (m < pr.m_total) {
      const int hw = pr.out_h * pr.out_w;
      b = m / hw;
      const int r = m - b * hw;
      ho = r / pr.out_w;
      wo = r - ho * pr.out_w;
      yb = b * pr.y_sn + ho * pr.y_sh + wo * pr.y_sw;
    }
    pix_b[tid] = b; pix_ho[tid] = ho; pix_wo[tid] = wo;
    pix_xbase[tid] = b * pr.x_sn;
    pix_ybase[tid] = yb;
  }

  const T* __restrict__ x = reinterpret_cast<const T*>(pr.x);
  const OT* __restrict__ off = reinterpret_cast<const OT*>(pr.offset);
  const OT* __restrict__ msk = reinterpret_cast<const OT*>(pr.mask);
  const T* __restrict__ wgt = reinterpret_cast<const T*>(p.w);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int kc = tid & (KC - 1);   // channel within slice (gather + weight load)
  const int prow = tid >> 4;       // 0..15
  const int tm = tid & 15;         // micro-tile row group  (pixels tm*4 .. tm*4+3)
  const int tn = tid >> 4;         // micro-tile col group  (channels tn*4 .. tn*4+3)

  const int dg_lo = gc0 / cpd, dg_hi = (gc1 - 1) / cpd;

#pragma unroll 1
  for (int tap = 0; tap < K; ++tap) {
    const int ti = tap / p.kw, tj = tap - ti * p.kw;
#pragma unroll 1
    for (int g = dg_lo; g <= dg_hi; ++g) {
      __syncthreads();   // previous users of meta_* / pix_* writers are done
      if (tid < BM) {
        Sample4 s;
#pragma unroll
        for (int i = 0; i < 4; ++i) { s.w[i] = 0.f; s.o[i] = 0; }
        if (pix_ybase[tid] >= 0) {
          const int b = pix_b[tid], ho = pix_ho[tid], wo = pix_wo[tid];
          float oy = 0.f, ox = 0.f, mk = 1.f;
          if (off != nullptr) {
            const int64_t o = b * pr.off_sn + (int64_t)(g * 2 * K + 2 * tap) * pr.off_sc + ho * pr.off_sh + wo * pr.off_sw;
            oy = to_f32(off[o]);
            ox = to_f32(off[o + pr.off_sc]);
          }
          if (msk != nullptr) {
            mk = to_f32(msk[b * pr.mask_sn + (int64_t)(g * K + tap) * pr.mask_sc + ho * pr.mask_sh + wo * pr.mask_sw]);
            if (p.flags & STM_DCN_MASK_SIGMOID) mk = sigmoidf_(mk);
          }
          const float h = (float)(ho * p.sh - p.ph + ti * p.dh) + oy;
          const float w = (float)(wo * p.sw - p.pw + tj * p.dw) + ox;
          s = make_sample(h, w, pr.in_h, pr.in_w, pr.x_sh, pr.x_sw, mk);
        }
        meta_w[tid] = make_float4(s.w[0], s.w[1], s.w[2], s.w[3]);
        meta_o[tid] = make_int4(s.o[0], s.o[1], s.o[2], s.o[3]);
      }
      __syncthreads();
      const int c_lo = max(gc0, g * cpd), c_hi = min(gc1, (g + 1) * cpd);
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += KC) {
        const int c = c0 + kc;
        const bool c_ok = c < c_hi;
        // ---- gather the A slice: As[kc][pixel] ----
#pragma unroll
        for (int pp = 0; pp < BM / 16; ++pp) {
          const int ml = prow + 16 * pp;
          float v = 0.f;
          if (c_ok) {
            const float4 w4 = meta_w[ml];
            const int4 o4 = meta_o[ml];
            const T* xb = x + pix_xbase[ml] + c;
            if (w4.x != 0.f) v = fmaf(w4.x, to_f32(xb[o4.x]), v);
            if (w4.y != 0.f) v = fmaf(w4.y, to_f32(xb[o4.y]), v);
            if (w4.z != 0.f) v = fmaf(w4.z, to_f32(xb[o4.z]), v);
            if (w4.w != 0.f) v = fmaf(w4.w, to_f32(xb[o4.w]), v);
          }
          As[kc][ml] = v;
        }
        // ---- weight slice: Ws[kc][n] = W[oc0+n][tap][c - gc0] ----
#pragma unroll
        for (int nn = 0; nn < BN / 16; ++nn) {
          const int n = prow + 16 * nn;
          const int oc = oc0 + n;
          float v = 0.f;
          if (c_ok && oc < oc_end) v = to_f32(wgt[((int64_t)oc * K + tap) * cpg + (c - gc0)]);
          Ws[kc][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tn * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }

  // ---- epilogue: bias, ReLU, store NHWC ----
  T* __restrict__ y = reinterpret_cast<T*>(pr.y);
  float* __restrict__ yf = reinterpret_cast<float*>(pr.y);
  const bool relu = (p.flags & STM_DCN_RELU) != 0;
  const bool out_f32 = (p.flags & STM_DCN_OUT_F32) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t yb = pix_ybase[tm * 4 + i];
    if (yb < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int oc = oc0 + tn * 4 + j;
      if (oc >= oc_end) continue;
      float v = acc[i][j];
      if (p.bias != nullptr) v += p.bias[oc];
      if (relu) v = fmaxf(v, 0.f);
      if (out_f32) yf[yb + oc] = v;
      else y[yb + oc] = from_f32<T>(v);
    }
  }
}

}  // namespace

int launch_dcn_simt(const DcnParams& p_in, int dtype, int offset_dtype, cudaStream_t stream) {
  DcnParams p = p_in;
  int tiles = 0;
  for (int i = 0; i < p.n_probs; ++i) {
    p.prob[i].tile_begin = tiles;
    tiles += (p.prob[i].m_total + BM - 1) / BM;
  }
  p.total_m_tiles = tiles;
  if (tiles == 0) return STM_OK;
  const int opg = p.out_c / p.groups;
  dim3 grid(tiles, (opg + BN - 1) / BN, p.groups);
  if (dtype == STM_F32 && offset_dtype == STM_F32)
    dcn_simt_kernel<float, float><<<grid, NT, 0, stream>>>(p);
  else if (dtype == STM_F32)
    dcn_simt_kernel<float, __nv_bfloat16><<<grid, NT, 0, stream>>>(p);
  else if (offset_dtype == STM_F32)
    dcn_simt_kernel<__nv_bfloat16, float><<<grid, NT, 0, stream>>>(p);
  else
    dcn_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, NT, 0, stream>>>(p);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace stm
