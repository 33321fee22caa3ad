// Candidate generation + cross-class fast NMS for a BATCH of frames, on the device and without a host round trip
// (the reference does this frame by frame in Python/torch with several device->host syncs:
//  generate_candidate, layers/functions/TF_utils.py:54-82, and Detect_TF.cc_fast_nms, layers/functions/detection_TF.py:85-134).
//
// One CTA per frame:
//   1. every prior: best foreground class probability and its class (conf[p, 1:].max), keep it if that probability
//      exceeds conf_thresh (TF_utils.py:68-71); score = probability x centerness (detection_TF.py:88-91); decode its box
//      from the regression and the prior (layers/box_utils.py:238-283, variances 0.1 / 0.2, point form);
//   2. the kept candidates are sorted by score, descending (ties: lower prior index first), with a bitonic sort in shared
//      memory, and cut to top_k (detection_TF.py:93-94);
//   3. fast NMS: candidate j survives iff no HIGHER-SCORING candidate i overlaps it by more than nms_thresh
//      (IoU = triu(jaccard, 1).max(0) <= thresh, detection_TF.py:100-118) — note that i need not survive itself;
//   4. survivors are written, in score order, to fixed-size outputs [frames, top_k] plus a per-frame count.
#include "common.cuh"

namespace stm {
namespace {

constexpr int NMS_THREADS = 1024;
constexpr int NMS_MAX_TOPK = 256;
constexpr int NMS_SORT_CAP = 16384;            // candidates per frame that take part in the sort (>= 15 345 priors of a 384x640 frame)

struct NmsArgs {
  const float* conf;        // [F, P, C] class probabilities (softmax applied)
  const float* loc;         // [F, P, 4]
  const float* centerness;  // [F, P] or null
  const float* priors;      // [P, 4] (cx, cy, w, h)
  int32_t* count;           // [F]
  int32_t* index;           // [F, top_k] prior index
  int32_t* cls;             // [F, top_k] class id (1-based: background is 0)
  float* score;             // [F, top_k]
  float* box;               // [F, top_k, 4] x1, y1, x2, y2
  int32_t P, C, top_k, pad_;
  float conf_thresh, nms_thresh;
};

__device__ __forceinline__ float4 decode_box(const float* loc, const float* pri) {
  // boxes = cat(priors[:, :2] + loc[:, :2] * 0.1 * priors[:, 2:], priors[:, 2:] * exp(loc[:, 2:] * 0.2)); to point form
  const float cx = pri[0] + loc[0] * 0.1f * pri[2], cy = pri[1] + loc[1] * 0.1f * pri[3];
  const float w = pri[2] * expf(loc[2] * 0.2f), h = pri[3] * expf(loc[3] * 0.2f);
  const float x1 = cx - w / 2.f, y1 = cy - h / 2.f;
  return make_float4(x1, y1, x1 + w, y1 + h);
}

// descending by score, ascending by prior index on ties; empty slots (idx < 0) last
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(NMS_THREADS) detect_nms_kernel(const __grid_constant__ NmsArgs a) {
  extern __shared__ unsigned char smem_raw[];
  float* s_score = reinterpret_cast<float*>(smem_raw);                      // [NMS_SORT_CAP]
  int* s_idx = reinterpret_cast<int*>(s_score + NMS_SORT_CAP);             // [NMS_SORT_CAP]
  __shared__ float4 s_box[NMS_MAX_TOPK];
  __shared__ int s_keep[NMS_MAX_TOPK];
  __shared__ int s_n;
  const int f = blockIdx.x, tid = threadIdx.x;
  const float* conf = a.conf + (size_t)f * a.P * a.C;
  const float* loc = a.loc + (size_t)f * a.P * 4;
  const float* ctr = a.centerness ? a.centerness + (size_t)f * a.P : nullptr;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // ---- 1. candidates ----
  for (int p = tid; p < a.P; p += NMS_THREADS) {
    const float* row = conf + (size_t)p * a.C;
    float best = row[1];
    for (int c = 2; c < a.C; ++c) best = fmaxf(best, row[c]);
    if (best > a.conf_thresh) {
      const int slot = atomicAdd(&s_n, 1);
      if (slot < NMS_SORT_CAP) {
        s_score[slot] = ctr ? best * ctr[p] : best;
        s_idx[slot] = p;
      }
    }
  }
  __syncthreads();
  const int n_cand = min(s_n, NMS_SORT_CAP);
  int n_sort = 1;
  while (n_sort < n_cand) n_sort <<= 1;
  for (int i = n_cand + tid; i < n_sort; i += NMS_THREADS) { s_score[i] = 0.f; s_idx[i] = -1; }
  __syncthreads();
  // ---- 2. bitonic sort (score desc, index asc) ----
  for (int k = 2; k <= n_sort; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_sort; i += NMS_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;               // this sub-sequence sorted "best first"
          const float si = s_score[i], sl = s_score[l];
          const int ii = s_idx[i], il = s_idx[l];
          const bool swap = up ? before(sl, il, si, ii) : before(si, ii, sl, il);
          if (swap) { s_score[i] = sl; s_score[l] = si; s_idx[i] = il; s_idx[l] = ii; }
        }
      }
      __syncthreads();
    }
  const int n = min(n_cand, a.top_k);
  // ---- 3. boxes of the top_k, fast NMS ----
  if (tid < n) s_box[tid] = decode_box(loc + (size_t)s_idx[tid] * 4, a.priors + (size_t)s_idx[tid] * 4);
  __syncthreads();
  if (tid < n) {
    const float4 b = s_box[tid];
    const float area_b = (b.z - b.x) * (b.w - b.y);
    float iou_max = 0.f;
    for (int i = 0; i < tid; ++i) {
      const float4 q = s_box[i];
      const float iw = fmaxf(fminf(q.z, b.z) - fmaxf(q.x, b.x), 0.f), ih = fmaxf(fminf(q.w, b.w) - fmaxf(q.y, b.y), 0.f);
      const float inter = iw * ih;
      const float uni = (q.z - q.x) * (q.w - q.y) + area_b - inter;
      iou_max = fmaxf(iou_max, inter / uni);
    }
    s_keep[tid] = iou_max <= a.nms_thresh ? 1 : 0;
  }
  __syncthreads();
  // ---- 4. compact in score order ----
  if (tid < n && s_keep[tid]) {
    int pos = 0;
    for (int i = 0; i < tid; ++i) pos += s_keep[i];
    const int p = s_idx[tid];
    const size_t o = (size_t)f * a.top_k + pos;
    const float* row = conf + (size_t)p * a.C;
    int best_c = 1;
    float best = row[1];
    for (int c = 2; c < a.C; ++c)
      if (row[c] > best) { best = row[c]; best_c = c; }        // first maximum, like torch.max
    a.index[o] = p;
    a.cls[o] = best_c;
    a.score[o] = s_score[tid];
    reinterpret_cast<float4*>(a.box)[o] = s_box[tid];
  }
  if (tid == 0) {
    int c = 0;
    for (int i = 0; i < n; ++i) c += s_keep[i];
    a.count[f] = c;
  }
}

SmemAttrCache g_nms_attr;

}  // namespace

int launch_detect_nms(const float* conf, const float* loc, const float* centerness, const float* priors, int frames, int P, int C,
                      int top_k, float conf_thresh, float nms_thresh, int32_t* count, int32_t* index, int32_t* cls, float* score,
                      float* box, cudaStream_t stream) {
  NmsArgs a;
  a.conf = conf; a.loc = loc; a.centerness = centerness; a.priors = priors;
  a.count = count; a.index = index; a.cls = cls; a.score = score; a.box = box;
  a.P = P; a.C = C; a.top_k = top_k; a.pad_ = 0;
  a.conf_thresh = conf_thresh; a.nms_thresh = nms_thresh;
  const int smem = NMS_SORT_CAP * 8;
  const int rc = ensure_dynamic_smem(detect_nms_kernel, smem, g_nms_attr);
  if (rc != STM_OK) return rc;
  detect_nms_kernel<<<frames, NMS_THREADS, smem, stream>>>(a);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace stm
