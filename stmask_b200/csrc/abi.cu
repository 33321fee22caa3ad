// extern "C" surface of libstmask_b200.so: argument validation, backend selection, launch.
// See include/stmask_b200.h for the contract and the reference interfaces each entry replaces.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <type_traits>

#include "common.cuh"

namespace stm {

static thread_local char g_err[512] = {0};
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void clear_error() { g_err[0] = 0; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int current_device_ordinal() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  return dev;
}

int device_sm_count() {
  static std::atomic<int> cache[64];
  const int dev = current_device_ordinal();
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    return 148;
  }
  if (dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

int pack_weight_ohwi(const void* src, int sd, void* dst, int dd, int O, int I, int K, cudaStream_t stream);
int transpose_batched(const void* src, int sd, void* dst, int dd, int n, int rows, int cols, cudaStream_t stream);
int ali_offsets(const void* shape, const int64_t* ss, int sd, void* off, const int64_t* os, int od, int B, int H, int W,
                int kh, int kw, cudaStream_t stream);

int ada_offsets(const void* shape, const int64_t* ss, int sd, const float* w, void* off, const int64_t* os, int od, int B,
                int H, int W, int OC, cudaStream_t stream);

static inline bool dtype_ok(int d) { return d == STM_F32 || d == STM_BF16; }
static inline size_t dsize(int d) { return d == STM_F32 ? 4 : 2; }
static inline int out_size(int in, int k, int s, int p, int d) { return (in + 2 * p - d * (k - 1) - 1) / s + 1; }

static int validate_conv(const StmDcnConv* c) {
  STM_CHECK_ARG(c != nullptr, "conv descriptor is null");
  STM_CHECK_ARG(c->in_c > 0 && c->out_c > 0, "in_c/out_c must be positive (got %d/%d)", c->in_c, c->out_c);
  STM_CHECK_ARG(c->kernel_h > 0 && c->kernel_w > 0, "kernel size must be positive");
  STM_CHECK_ARG(c->stride_h > 0 && c->stride_w > 0, "stride must be positive");
  STM_CHECK_ARG(c->dil_h > 0 && c->dil_w > 0, "dilation must be positive");
  STM_CHECK_ARG(c->pad_h >= 0 && c->pad_w >= 0, "padding must be non-negative");
  STM_CHECK_ARG(c->groups > 0 && c->in_c % c->groups == 0 && c->out_c % c->groups == 0,
                "in_c (%d) and out_c (%d) must be divisible by groups (%d)", c->in_c, c->out_c, c->groups);
  STM_CHECK_ARG(c->deform_groups > 0 && c->in_c % c->deform_groups == 0,
                "in_c (%d) must be divisible by deform_groups (%d)", c->in_c, c->deform_groups);
  STM_CHECK_ARG(dtype_ok(c->dtype) && dtype_ok(c->offset_dtype), "unknown dtype");
  STM_CHECK_ARG(c->backend >= STM_BACKEND_AUTO && c->backend <= STM_BACKEND_TCGEN05, "unknown backend %d", c->backend);
  return STM_OK;
}

static int validate_problems(const StmDcnConv* c, const StmDcnProblem* pr, int n) {
  STM_CHECK_ARG(pr != nullptr || n == 0, "problem array is null");
  STM_CHECK_ARG(n >= 0 && n <= STM_DCN_MAX_PROBLEMS, "n_probs %d outside [0, %d]", n, STM_DCN_MAX_PROBLEMS);
  for (int i = 0; i < n; ++i) {
    const StmDcnProblem& q = pr[i];
    STM_CHECK_ARG(q.batch >= 0 && q.in_h > 0 && q.in_w > 0, "problem %d: bad input size", i);
    const int oh = out_size(q.in_h, c->kernel_h, c->stride_h, c->pad_h, c->dil_h);
    const int ow = out_size(q.in_w, c->kernel_w, c->stride_w, c->pad_w, c->dil_w);
    STM_CHECK_ARG(oh > 0 && ow > 0, "problem %d: convolution output size is %dx%d", i, oh, ow);
    STM_CHECK_ARG(q.out_h == oh && q.out_w == ow, "problem %d: out size %dx%d, expected %dx%d", i, q.out_h, q.out_w, oh, ow);
    STM_CHECK_ARG((int64_t)q.batch * oh * ow < (1ll << 31), "problem %d: too many output pixels", i);
    if (q.batch == 0) continue;
    STM_CHECK_ARG(q.x != nullptr && q.y != nullptr, "problem %d: x/y pointer is null", i);
    STM_CHECK_ARG(q.x_stride_w >= c->in_c && ((c->flags & STM_DCN_OUT_PLANAR) ? q.y_stride_w >= 1 : q.y_stride_w >= c->out_c),
                  "problem %d: NHWC pixel stride smaller than C", i);
    STM_CHECK_ARG((int64_t)q.in_h * q.x_stride_h < (1ll << 31) && q.x_stride_h >= 0 && q.x_stride_w >= 0,
                  "problem %d: one image of x must span < 2^31 elements", i);
    if (!(c->flags & STM_DCN_ZERO_OFFSET)) STM_CHECK_ARG(q.offset != nullptr, "problem %d: offset pointer is null", i);
  }
  return STM_OK;
}

static void fill_params(const StmDcnConv* c, const StmDcnProblem* pr, int n, const void* w, const float* bias, DcnParams* p,
                        const float* fcb_w = nullptr) {
  memset(p, 0, sizeof(*p));
  p->fcb_w = fcb_w;
  p->n_probs = 0;
  for (int i = 0; i < n; ++i) {
    const StmDcnProblem& q = pr[i];
    if (q.batch == 0) continue;
    DcnProblemDev& d = p->prob[p->n_probs++];
    d.x = q.x; d.y = q.y;
    d.offset = (c->flags & STM_DCN_ZERO_OFFSET) ? nullptr : q.offset;
    d.mask = q.mask;
    d.batch = q.batch; d.in_h = q.in_h; d.in_w = q.in_w; d.out_h = q.out_h; d.out_w = q.out_w;
    d.m_total = q.batch * q.out_h * q.out_w;
    d.x_sn = q.x_stride_n; d.x_sh = q.x_stride_h; d.x_sw = q.x_stride_w;
    d.y_sn = q.y_stride_n; d.y_sh = q.y_stride_h; d.y_sw = q.y_stride_w;
    d.off_sn = q.off_stride_n; d.off_sc = q.off_stride_c; d.off_sh = q.off_stride_h; d.off_sw = q.off_stride_w;
    d.mask_sn = q.mask_stride_n; d.mask_sc = q.mask_stride_c; d.mask_sh = q.mask_stride_h; d.mask_sw = q.mask_stride_w;
  }
  p->in_c = c->in_c; p->out_c = c->out_c; p->kh = c->kernel_h; p->kw = c->kernel_w;
  p->sh = c->stride_h; p->sw = c->stride_w; p->ph = c->pad_h; p->pw = c->pad_w; p->dh = c->dil_h; p->dw = c->dil_w;
  p->groups = c->groups; p->dg = c->deform_groups; p->flags = c->flags;
  p->w = w; p->bias = bias; p->fcb_w = fcb_w;
}

static int pick_dcn_backend(const StmDcnConv* c, const StmDcnProblem* pr, int n) {
  const char* why = "";
  const bool tc = dcn_tc_supported(c, pr, n, &why);
  if (c->backend == STM_BACKEND_TCGEN05) {
    if (!tc) { set_error("tcgen05 deformable conv not available for this call: %s", why); return STM_ERR_UNSUPPORTED; }
    return STM_BACKEND_TCGEN05;
  }
  if (c->backend == STM_BACKEND_SIMT) return STM_BACKEND_SIMT;
  return tc ? STM_BACKEND_TCGEN05 : STM_BACKEND_SIMT;
}

static int validate_corr(const StmCorrDesc* d, const void* x1, const void* x2, const void* fa, const void* fb, const void* out) {
  STM_CHECK_ARG(d != nullptr, "correlation descriptor is null");
  STM_CHECK_ARG(d->batch >= 0 && d->h > 0 && d->w > 0 && d->c > 0, "bad correlation input size");
  STM_CHECK_ARG(d->patch > 0 && (d->patch & 1) == 1, "patch_size must be odd and positive (got %d)", d->patch);
  STM_CHECK_ARG(d->patch <= 31, "patch_size %d > 31 is not supported", d->patch);
  STM_CHECK_ARG(d->dilation_patch > 0, "dilation_patch must be positive");
  STM_CHECK_ARG(dtype_ok(d->dtype) && dtype_ok(d->out_dtype), "unknown dtype");
  STM_CHECK_ARG(d->backend >= STM_BACKEND_AUTO && d->backend <= STM_BACKEND_TCGEN05, "unknown backend %d", d->backend);
  if (d->batch == 0) return STM_OK;
  STM_CHECK_ARG(x1 && x2 && out, "x1/x2/out pointer is null");
  STM_CHECK_ARG(d->x1_stride_w >= d->c && d->x2_stride_w >= d->c, "NHWC pixel stride smaller than C");
  if (d->flags & STM_CORR_COPY_FEATS) {
    STM_CHECK_ARG(fa && fb && d->feat_c > 0 && dtype_ok(d->feat_dtype), "COPY_FEATS needs feat_a/feat_b/feat_c");
    STM_CHECK_ARG(d->feat_a_stride_w >= d->feat_c && d->feat_b_stride_w >= d->feat_c, "feat pixel stride smaller than feat_c");
    STM_CHECK_ARG(d->feat_c_offset == 0 || d->feat_c_offset >= d->patch * d->patch,
                  "feat_c_offset %d overlaps the %d correlation channels", d->feat_c_offset, d->patch * d->patch);
  }
  if (d->x1_index != nullptr || d->x2_index != nullptr) {
    STM_CHECK_ARG(d->x1_index != nullptr && d->x2_index != nullptr, "x1_index and x2_index must be given together");
    STM_CHECK_ARG(d->x1_frames > 0 && d->x2_frames > 0, "pair indexing needs x1_frames / x2_frames");
    STM_CHECK_ARG(d->alt_frames >= 0 && (d->alt_frames == 0 || d->x1_alt != nullptr), "alt_frames without x1_alt");
    if ((d->flags & STM_CORR_COPY_FEATS) && d->alt_frames > 0) STM_CHECK_ARG(d->feat_a_alt != nullptr, "alt_frames without feat_a_alt");
  }
  return STM_OK;
}

}  // namespace stm

using namespace stm;

extern "C" {

int stm_version(void) { return STM_ABI_VERSION; }
const char* stm_last_error(void) { return g_err; }
uint64_t stm_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int stm_device_supported(int32_t device) {
  clear_error();
  int major = 0, n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    (void)cudaGetLastError();
    return 0;
  }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

size_t stm_dcn_packed_weight_bytes(const StmDcnConv* c) {
  clear_error();
  if (validate_conv(c) != STM_OK) return 0;
  return (size_t)c->out_c * (c->in_c / c->groups) * c->kernel_h * c->kernel_w * dsize(c->dtype);
}

int stm_dcn_pack_weight(const StmDcnConv* c, const void* w_oihw, int32_t src_dtype, void* w_packed, void* stream) {
  clear_error();
  int rc = validate_conv(c);
  if (rc != STM_OK) return rc;
  STM_CHECK_ARG(w_oihw && w_packed, "weight pointer is null");
  STM_CHECK_ARG(dtype_ok(src_dtype), "unknown source dtype");
  return pack_weight_ohwi(w_oihw, src_dtype, w_packed, c->dtype, c->out_c, c->in_c / c->groups, c->kernel_h * c->kernel_w,
                          (cudaStream_t)stream);
}

size_t stm_deform_conv2d_workspace(const StmDcnConv* c, const StmDcnProblem* pr, int32_t n) {
  clear_error();
  if (validate_conv(c) != STM_OK || validate_problems(c, pr, n) != STM_OK) return 0;
  const int be = pick_dcn_backend(c, pr, n);
  return be == STM_BACKEND_TCGEN05 ? dcn_tc_workspace(c, pr, n) : 0;
}

int stm_deform_conv2d_backend(const StmDcnConv* c, const StmDcnProblem* pr, int32_t n) {
  clear_error();
  int rc = validate_conv(c);
  if (rc != STM_OK) return rc;
  rc = validate_problems(c, pr, n);
  if (rc != STM_OK) return rc;
  return pick_dcn_backend(c, pr, n);
}

int stm_deform_conv2d_variant(const StmDcnConv* c, const StmDcnProblem* pr, int32_t n, char* buf, size_t len) {
  clear_error();
  STM_CHECK_ARG(buf != nullptr && len > 0, "variant buffer is null");
  buf[0] = 0;
  int rc = validate_conv(c);
  if (rc != STM_OK) return rc;
  rc = validate_problems(c, pr, n);
  if (rc != STM_OK) return rc;
  // the plan is a function of the arguments (and the SM count) alone: it also answers without a driver
  const char* why = "";
  const bool tc = c->backend != STM_BACKEND_SIMT && dcn_tc_shape_supported(c, pr, n, &why);
  if (!tc && c->backend == STM_BACKEND_TCGEN05) { set_error("tcgen05 deformable conv not available for this call: %s", why); return STM_ERR_UNSUPPORTED; }
  if (!tc) { snprintf(buf, len, "simt"); return STM_OK; }
  DcnParams p;
  fill_params(c, pr, n, nullptr, nullptr, &p);
  if (p.n_probs == 0) { snprintf(buf, len, "empty"); return STM_OK; }
  return dcn_tc_variant(c, p, buf, len);
}

int stm_deform_conv2d_fwd(const StmDcnConv* c, const StmDcnProblem* pr, int32_t n, const void* w_packed, const float* bias,
                          void* workspace, size_t ws_bytes, void* stream) {
  clear_error();
  int rc = validate_conv(c);
  if (rc != STM_OK) return rc;
  rc = validate_problems(c, pr, n);
  if (rc != STM_OK) return rc;
  STM_CHECK_ARG(w_packed != nullptr, "packed weight pointer is null");
  STM_CHECK_ARG(!(c->flags & (STM_DCN_FCB_ADA | STM_DCN_FCB_ALI)), "STM_DCN_FCB_* flags belong to stm_deform_conv2d_fcb_fwd");
  const int be = pick_dcn_backend(c, pr, n);
  if (be < 0) return be;
  DcnParams p;
  fill_params(c, pr, n, w_packed, bias, &p);
  if (p.n_probs == 0) return STM_OK;
  if (be == STM_BACKEND_TCGEN05) return launch_dcn_tc(c, p, workspace, ws_bytes, (cudaStream_t)stream);
  if (c->flags & STM_DCN_OUT_PLANAR) { set_error("STM_DCN_OUT_PLANAR needs the tcgen05 backend"); return STM_ERR_UNSUPPORTED; }
  return launch_dcn_simt(p, c->dtype, c->offset_dtype, (cudaStream_t)stream);
}

int stm_deform_conv2d_fcb_fwd(const StmDcnConv* c, const StmDcnProblem* pr, int32_t n, const void* w_packed, const float* bias,
                              const float* fcb_weight, void* workspace, size_t ws_bytes, void* stream) {
  clear_error();
  int rc = validate_conv(c);
  if (rc != STM_OK) return rc;
  rc = validate_problems(c, pr, n);
  if (rc != STM_OK) return rc;
  STM_CHECK_ARG(w_packed != nullptr, "packed weight pointer is null");
  const int mode = c->flags & (STM_DCN_FCB_ADA | STM_DCN_FCB_ALI);
  STM_CHECK_ARG(mode == STM_DCN_FCB_ADA || mode == STM_DCN_FCB_ALI, "flags must carry exactly one of STM_DCN_FCB_ADA / STM_DCN_FCB_ALI");
  STM_CHECK_ARG(!(c->flags & STM_DCN_ZERO_OFFSET), "FCB offsets and STM_DCN_ZERO_OFFSET exclude each other");
  STM_CHECK_ARG(c->stride_h == 1 && c->stride_w == 1, "FCB calibration convs have stride 1 (box deltas are per output pixel)");
  if (mode == STM_DCN_FCB_ADA) STM_CHECK_ARG(fcb_weight != nullptr, "FCB(ada) needs the conv_offset weight");
  if (mode == STM_DCN_FCB_ALI)
    STM_CHECK_ARG(c->deform_groups == 1 && (c->kernel_h & 1) && (c->kernel_w & 1), "FCB(ali) offsets need deform_groups == 1 and an odd kernel");
  for (int i = 0; i < n; ++i) STM_CHECK_ARG(pr[i].mask == nullptr, "FCB calibration is DCNv1 (no mask)");
  const char* why = "";
  if (c->backend == STM_BACKEND_SIMT || !dcn_tc_supported(c, pr, n, &why)) {
    set_error("fused FCB offsets need the tcgen05 backend: %s", c->backend == STM_BACKEND_SIMT ? "SIMT backend requested" : why);
    return STM_ERR_UNSUPPORTED;
  }
  DcnParams p;
  fill_params(c, pr, n, w_packed, bias, &p, fcb_weight);
  if (p.n_probs == 0) return STM_OK;
  return launch_dcn_tc(c, p, workspace, ws_bytes, (cudaStream_t)stream);
}

int stm_fcb_ali_offsets(const void* shape, const int64_t ss[4], int32_t sd, void* offset, const int64_t os[4], int32_t od,
                        int32_t batch, int32_t h, int32_t w, int32_t kh, int32_t kw, void* stream) {
  clear_error();
  STM_CHECK_ARG(shape && offset && ss && os, "null pointer");
  STM_CHECK_ARG(batch >= 0 && h > 0 && w > 0 && kh > 0 && kw > 0, "bad size");
  STM_CHECK_ARG((kh & 1) && (kw & 1), "FCB(ali) offsets are defined for odd kernels (got %dx%d)", kh, kw);
  STM_CHECK_ARG(dtype_ok(sd) && dtype_ok(od), "unknown dtype");
  return ali_offsets(shape, ss, sd, offset, os, od, batch, h, w, kh, kw, (cudaStream_t)stream);
}

int stm_fcb_ada_offsets(const void* shape, const int64_t ss[4], int32_t sd, const float* weight, void* offset,
                        const int64_t os[4], int32_t od, int32_t batch, int32_t h, int32_t w, int32_t out_channels,
                        void* stream) {
  clear_error();
  STM_CHECK_ARG(shape && offset && ss && os && weight, "null pointer");
  STM_CHECK_ARG(batch >= 0 && h > 0 && w > 0, "bad size");
  STM_CHECK_ARG(out_channels > 0 && out_channels <= 2048, "out_channels %d outside (0, 2048]", out_channels);
  STM_CHECK_ARG(dtype_ok(sd) && dtype_ok(od), "unknown dtype");
  return ada_offsets(shape, ss, sd, weight, offset, os, od, batch, h, w, out_channels, (cudaStream_t)stream);
}

int stm_correlation_backend(const StmCorrDesc* d) {
  clear_error();
  STM_CHECK_ARG(d != nullptr, "correlation descriptor is null");
  const char* why = "";
  const bool tc = corr_tc_supported(*d, &why);
  if (d->backend == STM_BACKEND_TCGEN05) {
    if (!tc) { set_error("tcgen05 correlation not available for this call: %s", why); return STM_ERR_UNSUPPORTED; }
    return STM_BACKEND_TCGEN05;
  }
  if (d->backend == STM_BACKEND_SIMT) return STM_BACKEND_SIMT;
  return tc ? STM_BACKEND_TCGEN05 : STM_BACKEND_SIMT;
}

int stm_correlation_fwd(const StmCorrDesc* d, const void* x1, const void* x2, const void* fa, const void* fb, void* out,
                        void* stream) {
  clear_error();
  int rc = validate_corr(d, x1, x2, fa, fb, out);
  if (rc != STM_OK) return rc;
  if (d->batch == 0) return STM_OK;
  const int be = stm_correlation_backend(d);
  if (be < 0) return be;
  if (be == STM_BACKEND_TCGEN05) return launch_corr_tc(*d, x1, x2, fa, fb, out, (cudaStream_t)stream);
  if (d->x1_index != nullptr) { set_error("pair-indexed correlation needs the tcgen05 backend"); return STM_ERR_UNSUPPORTED; }
  return launch_corr_simt(*d, x1, x2, fa, fb, out, (cudaStream_t)stream);
}

int stm_correlation_multi_fwd(const StmCorrDesc* descs, const void* const* x1s, const void* const* x2s, void* const* outs,
                              int32_t n, void* stream) {
  clear_error();
  STM_CHECK_ARG(descs && x1s && x2s && outs, "null pointer");
  STM_CHECK_ARG(n >= 1 && n <= 8, "1..8 feature maps per launch (got %d)", n);
  const StmCorrDesc& d0 = descs[0];
  for (int i = 0; i < n; ++i) {
    const StmCorrDesc& d = descs[i];
    int rc = validate_corr(&d, x1s[i], x2s[i], nullptr, nullptr, outs[i]);
    if (rc != STM_OK) return rc;
    STM_CHECK_ARG(d.batch > 0, "feature map %d: empty batch", i);
    STM_CHECK_ARG(d.c == d0.c && d.patch == d0.patch && d.dilation_patch == d0.dilation_patch && d.dtype == d0.dtype &&
                  d.out_dtype == d0.out_dtype && d.flags == d0.flags && d.scale == d0.scale && d.leaky_slope == d0.leaky_slope &&
                  d.feat_c_offset == d0.feat_c_offset,
                  "feature map %d: C / patch / dilation / dtypes / flags / scale must match feature map 0", i);
    STM_CHECK_ARG(!(d.flags & STM_CORR_COPY_FEATS) && d.x1_index == nullptr, "grouped launches take no concat features and no pair indices");
    const char* why = "";
    if (!corr_tc_supported(d, &why)) { set_error("grouped correlation needs the tcgen05 backend: %s (feature map %d)", why, i); return STM_ERR_UNSUPPORTED; }
  }
  return launch_corr_tc_multi(descs, x1s, x2s, nullptr, nullptr, outs, n, (cudaStream_t)stream);
}

int stm_detect_fast_nms_fwd(const float* conf, const float* loc, const float* centerness, const float* priors, int32_t frames,
                            int32_t n_priors, int32_t n_classes, int32_t top_k, float conf_thresh, float nms_thresh,
                            int32_t* count, int32_t* index, int32_t* cls, float* score, float* box, void* stream) {
  clear_error();
  STM_CHECK_ARG(frames >= 0 && n_priors > 0 && n_classes >= 2, "bad size");
  STM_CHECK_ARG(top_k > 0 && top_k <= 256, "top_k %d outside (0, 256]", top_k);
  STM_CHECK_ARG(n_priors <= 16384, "at most 16384 priors per frame (got %d)", n_priors);
  if (frames == 0) return STM_OK;
  STM_CHECK_ARG(conf && loc && priors && count && index && cls && score && box, "null pointer");
  STM_CHECK_ARG((((uintptr_t)box) & 15) == 0, "box output must be 16-byte aligned");
  return launch_detect_nms(conf, loc, centerness, priors, frames, n_priors, n_classes, top_k, conf_thresh, nms_thresh, count, index,
                           cls, score, box, (cudaStream_t)stream);
}

int stm_mask_assembly_fwd(const float* proto, const float* coeff, const float* boxes, const int32_t* count, float* masks,
                          uint32_t* mask_bits, int32_t frames, int32_t h, int32_t w, int32_t k, int32_t max_n, void* stream) {
  clear_error();
  STM_CHECK_ARG(frames >= 0 && h > 0 && w > 0 && max_n >= 0, "bad size");
  STM_CHECK_ARG(k > 0 && k <= 64, "at most 64 prototypes (got %d)", k);
  STM_CHECK_ARG(max_n <= 65535 && frames <= 65535, "too many detections / frames for one launch");
  if (frames == 0 || max_n == 0) return STM_OK;
  STM_CHECK_ARG(proto && coeff && boxes && masks && mask_bits, "null pointer");
  return launch_mask_assembly(proto, coeff, boxes, count, masks, mask_bits, frames, h, w, k, max_n, (cudaStream_t)stream);
}

int stm_mask_iou_fwd(const uint32_t* bits_a, const uint32_t* bits_b, const int32_t* count_a, const int32_t* count_b, float* iou,
                     int32_t frames, int32_t max_a, int32_t max_b, int32_t words, void* stream) {
  clear_error();
  STM_CHECK_ARG(frames >= 0 && max_a >= 0 && max_b >= 0 && words > 0, "bad size");
  STM_CHECK_ARG(max_a <= 65535 && frames <= 65535, "too many masks / frames for one launch");
  if (frames == 0 || max_a == 0 || max_b == 0) return STM_OK;
  STM_CHECK_ARG(bits_a && bits_b && iou, "null pointer");
  return launch_mask_iou(bits_a, bits_b, count_a, count_b, iou, frames, max_a, max_b, words, (cudaStream_t)stream);
}

int stm_track_update_fwd(const StmTrackParams* p, const StmTrackState* st, const StmTrackDets* det, const float* mask_iou,
                         const uint8_t* is_first, int32_t* det_slot, uint8_t* keep, void* stream) {
  clear_error();
  STM_CHECK_ARG(p && st && det, "null descriptor");
  STM_CHECK_ARG(p->clips >= 0 && p->cap > 0 && p->max_det > 0, "bad clips / cap / max_det");
  STM_CHECK_ARG(p->k > 0 && p->e > 0 && p->words > 0 && p->hw >= 0, "bad k / e / words / hw");
  if (p->clips == 0) return STM_OK;
  STM_CHECK_ARG(st->n_obj && st->box && st->score && st->cls && st->coeff && st->track && st->tracked && st->mask_bits,
                "tracker state arrays must not be null (mask / centerness may)");
  STM_CHECK_ARG(det->box && det->score && det->cls && det->coeff && det->track && det->mask_bits, "detection arrays must not be null");
  STM_CHECK_ARG(mask_iou && det_slot && keep, "mask_iou / det_slot / keep pointer is null");
  return launch_track_update(*p, *st, *det, mask_iou, is_first, det_slot, keep, (cudaStream_t)stream);
}

int stm_roi_align_fwd(const StmRoiAlignDesc* d, const void* feat, const float* rois, void* out, void* stream) {
  clear_error();
  STM_CHECK_ARG(d != nullptr, "roi_align descriptor is null");
  STM_CHECK_ARG(d->batch > 0 && d->h > 0 && d->w > 0 && d->c > 0, "bad feature map size");
  STM_CHECK_ARG(d->n_rois >= 0 && d->pooled_h > 0 && d->pooled_w > 0, "bad roi count / output size");
  STM_CHECK_ARG(dtype_ok(d->dtype) && dtype_ok(d->out_dtype), "unknown dtype");
  STM_CHECK_ARG(d->feat_stride_w >= d->c, "NHWC pixel stride smaller than C");
  STM_CHECK_ARG((int64_t)d->n_rois * d->pooled_h * d->pooled_w < (1ll << 31), "too many output bins");
  if (d->n_rois == 0) return STM_OK;
  STM_CHECK_ARG(feat && rois && out, "feat/rois/out pointer is null");
  return launch_roi_align(*d, feat, rois, out, (cudaStream_t)stream);
}

int stm_pool_fc_fwd(const void* x, int32_t dtype, int32_t n, int32_t hw, int32_t c, int64_t x_stride_n, int64_t x_stride_p,
                    const float* weight, const float* bias, int32_t out_features, float* y, void* stream) {
  clear_error();
  STM_CHECK_ARG(dtype_ok(dtype), "unknown dtype");
  STM_CHECK_ARG(n >= 0 && hw > 0 && c > 0 && out_features > 0, "bad size");
  if (n == 0) return STM_OK;
  STM_CHECK_ARG(x && weight && y, "null pointer");
  STM_CHECK_ARG(x_stride_p >= c && x_stride_n >= 0, "pixel stride smaller than C");
  return launch_pool_fc(x, dtype, n, hw, c, x_stride_n, x_stride_p, weight, bias, out_features, y, (cudaStream_t)stream);
}

int stm_nchw_to_nhwc(const void* src, int32_t sd, void* dst, int32_t dd, int32_t n, int32_t c, int32_t h, int32_t w,
                     void* stream) {
  clear_error();
  STM_CHECK_ARG(src && dst, "null pointer");
  STM_CHECK_ARG(n >= 0 && c > 0 && h > 0 && w > 0, "bad size");
  STM_CHECK_ARG(dtype_ok(sd) && dtype_ok(dd), "unknown dtype");
  return transpose_batched(src, sd, dst, dd, n, c, h * w, (cudaStream_t)stream);   // [c][hw] -> [hw][c]
}

int stm_nhwc_to_nchw(const void* src, int32_t sd, void* dst, int32_t dd, int32_t n, int32_t c, int32_t h, int32_t w,
                     void* stream) {
  clear_error();
  STM_CHECK_ARG(src && dst, "null pointer");
  STM_CHECK_ARG(n >= 0 && c > 0 && h > 0 && w > 0, "bad size");
  STM_CHECK_ARG(dtype_ok(sd) && dtype_ok(dd), "unknown dtype");
  return transpose_batched(src, sd, dst, dd, n, h * w, c, (cudaStream_t)stream);   // [hw][c] -> [c][hw]
}

}  // extern "C"
