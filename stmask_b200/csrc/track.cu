// Device-side tracker state machine: the matching step of Track_TF.track (reference layers/functions/track_TF.py:52-181)
// for a batch of independent clips, one CTA per clip, no host round trip.
//
// The reference keeps `prev_candidate` as Python dicts of tensors that grow with torch.cat and decides every
// assignment in a Python loop over `match_ids` (track_TF.py:132-156) — a device->host sync per detection.  Here the
// state is a set of fixed-capacity device arrays (StmTrackState) and one launch per frame does, per clip:
//   1. age the tracked objects (tracked_mask + 1, track_TF.py:64,103),
//   2. comprehensive matching scores of every detection against {new object} + every tracked object
//      (compute_comp_scores, TF_utils.py:98-123: (cos + 1) / 2 of the track embeddings, + c0 * score + c1 * mask IoU
//      + c2 * box IoU + c3 * same-label, dummy column 0 with IoU 0.3) and their first arg-max (torch.max),
//   3. the SEQUENTIAL assignment loop, in detection order, by one thread (new object -> append; matched -> the
//      detection with the highest score wins the object, ties to the earlier one),
//   4. the row copies the assignments imply (box, score, class, mask coefficients, track embedding, centerness,
//      mask bit plane, soft mask), by the whole CTA,
//   5. the output filter (tracked_mask <= max_age, mask area > 1 pixel, score > conf_thresh; track_TF.py:158-165).
// What CandidateShift does to the state before this step (shifted boxes / coefficients / masks, score x 0.95) is the
// caller's: the correlation, RoIAlign, TemporalNet and mask-assembly kernels of this library.
#include "common.cuh"

namespace stm {
namespace {

constexpr int TRACK_THREADS = 256;
constexpr int TRACK_MAX_CAP = 256;
constexpr int TRACK_MAX_DET = 256;

struct TrackArgs {
  StmTrackParams p;
  StmTrackState st;
  StmTrackDets det;
  const float* mask_iou;
  const uint8_t* is_first;
  int32_t* det_slot;
  uint8_t* keep;
};

__device__ __forceinline__ float box_iou(const float* a, const float* b) {
  // jaccard (box_utils.py:60-88): no clamping of the boxes themselves, only of the intersection extents
  const float iw = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f);
  const float ih = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
  const float inter = iw * ih;
  const float area_a = (a[2] - a[0]) * (a[3] - a[1]);
  const float area_b = (b[2] - b[0]) * (b[3] - b[1]);
  return inter / (area_a + area_b - inter);
}

__global__ void __launch_bounds__(TRACK_THREADS) track_update_kernel(const TrackArgs a) {
  __shared__ int s_match[TRACK_MAX_DET];       // arg-max column per detection (0 = new object)
  __shared__ int s_slot[TRACK_MAX_DET];        // state slot a detection's rows go to, -1: none
  __shared__ int s_src[TRACK_MAX_CAP];         // detection whose rows a slot takes this frame, -1: keeps its own
  __shared__ float s_best[TRACK_MAX_CAP];
  __shared__ int s_besti[TRACK_MAX_CAP];
  __shared__ int s_nout;
  extern __shared__ float s_vec[];             // [warps][e]: the detection's track embedding

  const StmTrackParams& p = a.p;
  const int clip = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cap = p.cap, maxd = p.max_det;
  int n_det = a.det.count != nullptr ? a.det.count[clip] : maxd;
  n_det = max(0, min(n_det, maxd));
  int n_prev = a.st.n_obj[clip];
  if (a.is_first != nullptr && a.is_first[clip]) n_prev = 0;          // track_TF.py:54-56: a new video forgets the state
  n_prev = max(0, min(n_prev, cap));

  const int64_t so = (int64_t)clip * cap;      // first state row of this clip
  const int64_t dofs = (int64_t)clip * maxd;   // first detection row

  for (int i = tid; i < maxd; i += TRACK_THREADS) s_slot[i] = -1;
  for (int j = tid; j < cap; j += TRACK_THREADS) { s_src[j] = -1; s_best[j] = -1.f; s_besti[j] = -1; }
  if (tid == 0) s_nout = n_prev;
  __syncthreads();

  if (n_det > 0 && n_prev == 0) {
    // ---- first frame of a clip (or nothing tracked yet): the detections become the state (track_TF.py:88-94) ----
    for (int i = tid; i < n_det && i < cap; i += TRACK_THREADS) { s_slot[i] = i; s_src[i] = i; }
    if (tid == 0) s_nout = min(n_det, cap);
  } else if (n_prev > 0) {
    // ---- every tracked object ages by one frame (track_TF.py:64,103) ----
    for (int j = tid; j < n_prev; j += TRACK_THREADS) a.st.tracked[so + j] += 1;
    if (n_det > 0) {
      // ---- comprehensive scores + first arg-max, one warp per detection ----
      float* vec = s_vec + warp * p.e;
      for (int i = warp; i < n_det; i += TRACK_THREADS / 32) {
        const float* dt = a.det.track + (dofs + i) * p.e;
        for (int c = lane; c < p.e; c += 32) vec[c] = dt[c];
        __syncwarp();
        const float dscore = a.det.score[dofs + i];
        const int dcls = a.det.cls[dofs + i];
        float dbox[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) dbox[c] = a.det.box[(dofs + i) * 4 + c];
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int col = lane; col <= n_prev; col += 32) {
          float ll, miou, biou, same;
          if (col == 0) {
            ll = 0.5f; miou = p.bbox_dummy_iou; biou = p.bbox_dummy_iou; same = 1.f;
          } else {
            const int j = col - 1;
            const float* pt = a.st.track + (so + j) * p.e;
            float dot = 0.f;
            for (int c = 0; c < p.e; ++c) dot = fmaf(vec[c], pt[c], dot);
            ll = (dot + 1.f) / 2.f;
            miou = a.mask_iou[(dofs + i) * cap + j];
            biou = box_iou(dbox, a.st.box + (so + j) * 4);
            same = a.st.cls[so + j] == dcls ? 1.f : 0.f;
          }
          // TF_utils.py:119-123, left to right
          float comp = ll + p.match_coeff[0] * dscore;
          comp = comp + p.match_coeff[1] * miou;
          comp = comp + p.match_coeff[2] * biou;
          comp = comp + p.match_coeff[3] * same;
          if (comp > best) { best = comp; besti = col; }             // strided columns ascend: keeps the first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) s_match[i] = besti == 0x7fffffff ? 0 : besti;
        __syncwarp();
      }
      __syncthreads();
      // ---- the reference's sequential assignment (track_TF.py:132-156), in detection order ----
      if (tid == 0) {
        int n_cur = n_prev;
        for (int i = 0; i < n_det; ++i) {
          const int m = s_match[i];
          if (m == 0) {
            if (n_cur < cap) { s_slot[i] = n_cur; s_src[n_cur] = i; ++n_cur; }      // a new object (dropped when the state is full)
          } else {
            const int obj = m - 1;
            const float sc = a.det.score[dofs + i];
            if (sc > s_best[obj]) {
              if (s_besti[obj] != -1) s_slot[s_besti[obj]] = -1;
              s_slot[i] = obj;
              s_src[obj] = i;
              s_best[obj] = sc;
              s_besti[obj] = i;
            }
          }
        }
        s_nout = n_cur;
      }
    }
  }
  __syncthreads();
  const int n_out = s_nout;

  // ---- row copies: slot <- its detection ----
  for (int s = warp; s < n_out; s += TRACK_THREADS / 32) {
    const int i = s_src[s];
    if (i < 0) continue;
    const int64_t sr = so + s, dr = dofs + i;
    if (lane < 4) a.st.box[sr * 4 + lane] = a.det.box[dr * 4 + lane];
    if (lane == 4) a.st.score[sr] = a.det.score[dr];
    if (lane == 5) a.st.cls[sr] = a.det.cls[dr];
    if (lane == 6) a.st.tracked[sr] = 0;
    if (lane == 7 && a.st.centerness != nullptr) a.st.centerness[sr] = a.det.centerness != nullptr ? a.det.centerness[dr] : 0.f;
    for (int c = lane; c < p.k; c += 32) a.st.coeff[sr * p.k + c] = a.det.coeff[dr * p.k + c];
    for (int c = lane; c < p.e; c += 32) a.st.track[sr * p.e + c] = a.det.track[dr * p.e + c];
    for (int c = lane; c < p.words; c += 32) a.st.mask_bits[sr * p.words + c] = a.det.mask_bits[dr * p.words + c];
  }
  if (a.st.mask != nullptr && a.det.mask != nullptr) {
    const bool vec4 = (p.hw & 3) == 0 && (((uintptr_t)a.st.mask | (uintptr_t)a.det.mask) & 15) == 0;
    for (int s = 0; s < n_out; ++s) {
      const int i = s_src[s];
      if (i < 0) continue;
      const float* src = a.det.mask + (dofs + i) * p.hw;
      float* dst = a.st.mask + (so + s) * p.hw;
      if (vec4) {
        for (int c = tid; c < p.hw / 4; c += TRACK_THREADS) reinterpret_cast<float4*>(dst)[c] = reinterpret_cast<const float4*>(src)[c];
      } else {
        for (int c = tid; c < p.hw; c += TRACK_THREADS) dst[c] = src[c];
      }
    }
  }
  __syncthreads();

  // ---- outputs ----
  for (int i = tid; i < maxd; i += TRACK_THREADS) a.det_slot[dofs + i] = s_slot[i];
  for (int s = warp; s < cap; s += TRACK_THREADS / 32) {
    bool k = false;
    if (s < n_out) {
      int area = 0;
      for (int c = lane; c < p.words; c += 32) area += __popc(a.st.mask_bits[(so + s) * p.words + c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
      k = a.st.tracked[so + s] <= p.max_age && area > 1 && a.st.score[so + s] > p.conf_thresh;
    }
    if (lane == 0) a.keep[so + s] = k ? 1 : 0;
  }
  if (tid == 0) a.st.n_obj[clip] = n_out;
}

}  // namespace

int launch_track_update(const StmTrackParams& p, const StmTrackState& st, const StmTrackDets& det, const float* mask_iou,
                        const uint8_t* is_first, int32_t* det_slot, uint8_t* keep, cudaStream_t stream) {
  if (p.cap > TRACK_MAX_CAP || p.max_det > TRACK_MAX_DET) {
    set_error("tracker capacity %d / max_det %d above the kernel's limit (%d / %d)", p.cap, p.max_det, TRACK_MAX_CAP, TRACK_MAX_DET);
    return STM_ERR_UNSUPPORTED;
  }
  if (p.clips == 0) return STM_OK;
  TrackArgs a;
  a.p = p; a.st = st; a.det = det; a.mask_iou = mask_iou; a.is_first = is_first; a.det_slot = det_slot; a.keep = keep;
  const size_t smem = (size_t)(TRACK_THREADS / 32) * p.e * sizeof(float);
  if (smem > 40 * 1024) { set_error("track embedding dimension %d too large", p.e); return STM_ERR_UNSUPPORTED; }
  track_update_kernel<<<p.clips, TRACK_THREADS, smem, stream>>>(a);
  count_launch();
  STM_CUDA_OK(cudaGetLastError());
  return STM_OK;
}

}  // namespace stm
