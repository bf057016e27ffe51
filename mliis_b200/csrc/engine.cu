// mliis_b200 engine: the C ABI of include/mliis_b200.h on top of the kernels in this directory.
// One inner step = forward (train-mode BN) -> fused loss -> backward -> BN EMA -> optimizer, all on one
// CUDA stream with every tensor resident in the slot's workspace; nothing crosses to the host.
//
// Reference call sites replaced: see include/mliis_b200.h.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mliis_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "plan.h"

using namespace mliis;

namespace mliis {
unsigned long long g_kernel_launches = 0;
ZGroup& zgroup() {
  thread_local ZGroup g{1, 0};
  return g;
}
int partition_nz() {
  const char* e = getenv("MLIIS_GROUP_CANONICAL");
  return (e && atoi(e) != 0) ? 1 : zgroup().nz;
}
// the group set by mliis_kernel_group for the calling thread (per-kernel entry points and mliis_train_step)
static thread_local ZGroup t_kernel_group{1, 0};
bool skip_launch(const char* launcher) {
  static const char* env = getenv("MLIIS_SKIP");
  if (!env || !*env) return false;
  const char* p = env;
  while (*p) {
    const char* e = strchr(p, ',');
    const size_t n = e ? (size_t)(e - p) : strlen(p);
    if (n > 0) {
      std::string tok(p, n);
      if (tok.back() == '$') {        // "name$" = exact launcher name ("tc_conv$" does not match tc_conv3)
        tok.pop_back();
        if (tok == launcher) return true;
      } else if (strstr(launcher, tok.c_str())) {
        return true;
      }
    }
    p += n + (e ? 1 : 0);
  }
  return false;
}
}

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

struct Slot {
  float* state = nullptr;
  float* ws = nullptr;
  int fwd_batch = 0;
  bool fwd_training = false;
  const float* fwd_images = nullptr;
  const int32_t* fwd_index = nullptr;
  const float* fwd_drop_mask = nullptr;
  const int32_t* class_ids = nullptr;   // multi-class mode: per-example foreground class (mliis_set_class_ids)
  bool grads_zeroed = false;
  bool folded[4] = {false, false, false, false};   // per RSD module: forward folded the pooled branch (k_pool.cu)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  unsigned long long graph_launches = 0;   // kernels inside the captured task graph
};

struct Tab {
  int32_t *lo = nullptr, *hi = nullptr, *g_lo = nullptr, *g_hi = nullptr;
  float* lerp = nullptr;
  ResizeTab rt() const { return ResizeTab{lo, hi, lerp, g_lo, g_hi}; }
};

}  // namespace

// NCCL is bound at run time (dlopen): the library has no link-time dependency on it and single-GPU users never load
// it.  Only the five entry points of the meta-update exchange are used.
struct NcclId { char internal[128]; };     // ncclUniqueId
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value: */ NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace

struct mliis_ctx {
  void* nccl_comm = nullptr;     // ncclComm_t of the meta-update exchange (mliis_comm_init); null = single process
  int comm_rank = 0, comm_world = 1;
  bool group_fallback = false;   // a task-batched call reached a kernel that only serves one slot (fp32 FFMA GEMMs)
  mliis_config cfg;
  Plan plan;
  int device = -1;
  std::vector<Slot> slots;
  int32_t* d_gamma_idx = nullptr;
  int32_t* d_beta_idx = nullptr;
  TcPrepJob* d_prep_jobs = nullptr;   // device copy of plan.prep_jobs with the per-layer TF32 split resolved
  std::vector<Tab> tabs;
  std::vector<void*> owned;
  float keep[16];
};

namespace {

// TF ResizeBilinear(align_corners=True) tables, computed in float32 exactly like the oracle [TF-ext]
void host_tables(int n_in, int n_out, std::vector<int32_t>& lo, std::vector<int32_t>& hi, std::vector<float>& lerp,
                 std::vector<int32_t>& g_lo, std::vector<int32_t>& g_hi) {
  lo.resize(n_out); hi.resize(n_out); lerp.resize(n_out);
  g_lo.assign(n_in, n_out); g_hi.assign(n_in, -1);
  const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
  for (int i = 0; i < n_out; ++i) {
    const float src = (float)i * scale;
    const int l = (int)floorf(src);
    int h = (int)ceilf(src);
    if (h > n_in - 1) h = n_in - 1;
    lo[i] = l; hi[i] = h; lerp[i] = src - (float)l;
    for (int j : {l, h}) {
      if (i < g_lo[j]) g_lo[j] = i;
      if (i > g_hi[j]) g_hi[j] = i;
    }
  }
}

template <typename T>
T* upload(mliis_ctx* c, const std::vector<T>& v) {
  T* d = nullptr;
  if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  if (c) c->owned.push_back(d);
  return d;
}

bool make_tab(mliis_ctx* c, int n_in, int n_out, Tab* t) {
  std::vector<int32_t> lo, hi, glo, ghi;
  std::vector<float> lerp;
  host_tables(n_in, n_out, lo, hi, lerp, glo, ghi);
  t->lo = upload(c, lo); t->hi = upload(c, hi); t->lerp = upload(c, lerp);
  t->g_lo = upload(c, glo); t->g_hi = upload(c, ghi);
  return t->lo && t->hi && t->lerp && t->g_lo && t->g_hi;
}

int check_cuda(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MLIIS_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
  return MLIIS_OK;
}

__global__ void dcs_kernel(const float* __restrict__ mask, float* __restrict__ dcs, int n_dc, int B, int maxB,
                           float k0, float k1, float k2, float k3, float k4, float k5, float k6, float k7, long long zs) {
  { const size_t zo = (size_t)blockIdx.z * zs; mask = zp(mask, zo); dcs += zo; }
  const float keep[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dc * B) return;
  int d = i / B, b = i - d * B;
  // utils.py:157-170: tf.div(inputs, keep_prob) * binary_tensor
  dcs[d * maxB + b] = (mask ? mask[i] : 1.f) / keep[d];
}

__global__ void set_scalar_kernel(float* p, float v, long long zs) { p[(size_t)blockIdx.z * zs] = v; }

struct Run {
  mliis_ctx* c;
  Slot* sl;
  const Plan& p;
  int B;
  cudaStream_t st;
  float* theta;
  float* mm;
  float* mv;
  float* adam_v;
  float* powers;
  float* W(int64_t off) const { return sl->ws + off; }
  float* T(int64_t off) const { return theta + off; }
  float* G(int64_t off) const { return sl->ws + p.grads + off; }
  Run(mliis_ctx* ctx, int slot, int batch, cudaStream_t s)
      : c(ctx), sl(&ctx->slots[slot]), p(ctx->plan), B(batch), st(s) {
    theta = sl->state;
    mm = sl->state + p.n_theta;
    mv = mm + p.n_bn_ch;
    adam_v = mv + p.n_bn_ch;
    powers = adam_v + p.n_theta;
  }
  // the self-resetting ticket word of the clustered reductions: the last floats of the partials area (slack)
  unsigned* ticket() const { return reinterpret_cast<unsigned*>(W(p.partials + p.partials_len - 16)); }
  float* bn_a(const BnRef& r) const { return W(p.bn_a + r.off); }
  float* bn_b(const BnRef& r) const { return W(p.bn_b + r.off); }
  float* bn_mean(const BnRef& r) const { return W(p.bn_mean + r.off); }
  float* bn_rstd(const BnRef& r) const { return W(p.bn_rstd + r.off); }

  // train-mode batch statistics of one BN layer (+ EMA of the moving statistics)
  void bn_train(const BnRef& r, const float* x, int ld, int M, bool pre_swish) const {
    bn_stats_finalize(x, ld, M, r.C, pre_swish, W(p.partials), ticket(), T(r.gamma), T(r.beta), mm + r.off, mv + r.off, 1,
                      r.fused, bn_mean(r), bn_rstd(r), bn_a(r), bn_b(r), st);
  }
};

GemmA plainA(const float* ptr, int ld) {
  GemmA a{};
  a.ptr = ptr; a.ld = ld;
  return a;
}
GemmA convA(const float* ptr, int ld, int H, int W, int C, int dil) {
  GemmA a{};
  a.ptr = ptr; a.ld = ld; a.conv = 1; a.H = H; a.W = W; a.C = C; a.dil = dil;
  return a;
}


// Dense contraction dispatch: tcgen05 (TF32) when the ctx asks for it and the shape is supported, else fp32 FFMA.
// w_hwio is the TF-layout kernel [taps][Cin][Cout]; scratch receives the re-laid-out tensor-core operand.
struct Dense {
  const Run& r;
  bool decoder = false;   // MLIIS_GEMM_TF32: single-pass TF32 for the decoder convs only; the backbone stays 3xTF32
  bool tc() const { return r.c->cfg.gemm_mode != MLIIS_GEMM_FP32; }
  int split() const { return (r.c->cfg.gemm_mode == MLIIS_GEMM_TF32 && decoder) ? 1 : 3; }
  // operand prepared by tc_prep_all at the start of the step (nullptr for a weight outside the plan's job table)
  float* prepared(const float* w_hwio, int dgrad, int Cin) const {
    if (!r.c->d_prep_jobs) return nullptr;
    const int64_t off = w_hwio - r.theta;
    for (const Plan::PrepJob& j : r.p.prep_jobs)
      if (j.w_off == off && j.dgrad == dgrad && j.Ci == Cin) return r.W(r.p.wcache + j.dst);
    return nullptr;
  }
  // Cs: input channels per tap of the stored kernel when the operand only uses its first Cin (0 = Cin)
  float* operand(const float* w_hwio, int taps, int Cin, int Cout, int dgrad, int Cs = 0) const {
    if (float* wt = prepared(w_hwio, dgrad, Cin)) return wt;
    float* wt = r.W(r.p.wT);
    tc_prep_weights(w_hwio, wt, taps, Cin, Cout, dgrad, split(), r.st, Cs);
    return wt;
  }
  // Folded pooled branch (k_pool.cu): is the 3x3 layer served by the kernel that adds the border-class bias?
  bool fold() const {
    static int off = -1;
    if (off < 0) { const char* e = getenv("MLIIS_NO_FOLD"); off = e ? atoi(e) : 0; }
    return tc() && !off;
  }
  // out = conv3x3(A[:, :Cin] ; first Cin of the Cs input channels of w) + bias + bias9[img][border class]
  bool fwd_fold(const float* A, int lda, int H, int W, int Cin, int Cs, const float* w_hwio, const float* bias,
                const float* bias9, float* out, int ldc, int Cout, int M) const {
    if (!tc_supported(1, W, Cin, Cout)) return false;
    float* wt = operand(w_hwio, 9, Cin, Cout, 0, Cs);
    return tc_conv(A, lda, wt, bias, out, ldc, 1, M, r.B, H, W, Cin, 9, 1, Cout, 0, split(), r.st, nullptr, nullptr,
                   nullptr, 0, bias9);
  }
  // forward: out[M, Cout] = conv(A[.., Cin]) + bias
  void fwd(const float* A, int lda, int conv, int H, int W, int Cin, int dil, const float* w_hwio, const float* bias,
           float* out, int ldc, int Cout, int M, int HW) const {
    const int taps = conv ? 9 : 1;
    if (tc() && tc_supported(conv, W, Cin, Cout)) {
      float* wt = operand(w_hwio, taps, Cin, Cout, 0);
      if (tc_conv(A, lda, wt, bias, out, ldc, conv, M, r.B, H, W, Cin, taps, dil, Cout, 0, split(), r.st)) return;
    }
    if (MLIIS_NZ > 1) r.c->group_fallback = true;
    GemmA a = conv ? convA(A, lda, H, W, Cin, dil) : plainA(A, lda);
    gemm_nn(a, w_hwio, bias, out, ldc, M, taps * Cin, Cout, HW, 0, r.st);
  }
  // MBConv project conv: out[M, Cout] = (swish(pa*A+pb) * gate[img]) * W
  void fwd_pro(const float* A, int lda, int Cin, const float* w_hwio, float* out, int ldc, int Cout, int M, int HW,
               const float* pa, const float* pb, const float* gate) const {
    if (tc() && tc_supported(0, 0, Cin, Cout)) {
      float* wt = operand(w_hwio, 1, Cin, Cout, 0);
      if (tc_conv(A, lda, wt, nullptr, out, ldc, 0, M, r.B, 1, 1, Cin, 1, 1, Cout, 0, split(), r.st, pa, pb, gate, HW))
        return;
    }
    if (MLIIS_NZ > 1) r.c->group_fallback = true;
    GemmA a = plainA(A, lda);
    a.pa = pa; a.pb = pb; a.gate = gate;
    gemm_nn(a, w_hwio, nullptr, out, ldc, M, Cin, Cout, HW, 0, r.st);
  }
  // wgrad: dW[taps*Cin, Cout] = sum_pixels A^T G (+ dbias = column sums of G).  `swap`: compute dW^T with the roles
  // of A and G exchanged (MBConv expand conv: Cout = 6*Cin > 256 would not fit one N tile), then transpose.
  void wgrad(const float* A, int lda, int conv, int H, int W, int Cin, int dil, const float* G, int ldg, int Cout,
             float* dW, float* dbias, int M, int HW, const float* pa = nullptr, const float* pb = nullptr,
             const float* gate = nullptr, bool swap = false, int dw_tap_stride = 0) const {
    const int taps = conv ? 9 : 1;
    float* scratch = r.W(r.p.tn_scratch);
    bool done = false;
    if (tc()) {
      if (!swap && tc_wgrad_supported(conv, W, Cin, Cout)) {
        done = tc_wgrad(A, lda, G, ldg, dW, scratch, conv, M, r.B, H, W, Cin, taps, dil, Cout, split(), r.st, pa, pb, gate, HW,
                        dw_tap_stride);
      } else if (swap && !conv && !pa && tc_wgrad_supported(0, W, Cout, Cin)) {
        float* tmp = r.W(r.p.wT);      // dW^T [Cout][Cin]
        done = tc_wgrad(G, ldg, A, lda, tmp, scratch, 0, M, r.B, H, W, Cout, 1, 1, Cin, split(), r.st);
        if (done) transpose_w(tmp, dW, Cout, Cin, r.st);
      }
      if (done && dbias) img_colsum(G, ldg, 1, M, Cout, 1.f, r.W(r.p.partials), dbias, Cout, r.st);
    }
    if (!done) {
      if (MLIIS_NZ > 1) r.c->group_fallback = true;
      GemmA a = conv ? convA(A, lda, H, W, Cin, dil) : plainA(A, lda);
      a.pa = pa; a.pb = pb; a.gate = gate;
      gemm_tn(a, G, ldg, dW, dbias, scratch, M, taps * Cin, Cout, HW, r.st);
    }
  }
  // dgrad: dA[M, Cin] (+)= conv^T(G[.., Cout])
  void dgrad(const float* G, int ldg, int conv, int H, int W, int Cin, int dil, const float* w_hwio, float* dA, int ldd,
             int Cout, int M, int HW, int accumulate, int Cs = 0) const {
    const int taps = conv ? 9 : 1;
    float* wt = r.W(r.p.wT);
    if (tc() && tc_supported(conv, W, Cout, Cin)) {
      float* wtc = operand(w_hwio, taps, Cin, Cout, 1, Cs);
      if (tc_conv(G, ldg, wtc, nullptr, dA, ldd, conv, M, r.B, H, W, Cout, taps, dil, Cin, accumulate, split(), r.st)) return;
    }
    if (MLIIS_NZ > 1) r.c->group_fallback = true;
    if (conv) {
      flip_transpose_w3x3(w_hwio, wt, Cin, Cout, r.st);
      gemm_nn(convA(G, ldg, H, W, Cout, dil), wt, nullptr, dA, ldd, M, 9 * Cout, Cin, HW, accumulate, r.st);
    } else {
      transpose_w(w_hwio, wt, Cin, Cout, r.st);
      gemm_nn(plainA(G, ldg), wt, nullptr, dA, ldd, M, Cout, Cin, HW, accumulate, r.st);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
void run_forward(const Run& r, const float* images, const int32_t* index, bool training, const float* dc_mask,
                 const float* drop_mask, uint64_t seed, const uint64_t* seed_dev = nullptr) {
  const Plan& p = r.p;
  const int B = r.B;
  cudaStream_t st = r.st;
  if (r.c->d_prep_jobs) tc_prep_all(r.theta, r.W(p.wcache), r.c->d_prep_jobs, (int)p.prep_jobs.size(), st);
  if (!training)
    bn_eval_coeffs(r.theta, r.c->d_gamma_idx, r.c->d_beta_idx, r.mm, r.mv, p.n_bn_ch, r.W(p.bn_a), r.W(p.bn_b), st);
  if (training && p.n_dc > 0) {
    const float* k = r.c->keep;
    MLIIS_COUNT(), dcs_kernel<<<dim3(cdiv(p.n_dc * B, 128), 1, MLIIS_NZ), 128, 0, st>>>(dc_mask, r.W(p.dcs), p.n_dc, B, p.maxB, k[0], k[1],
                                                                         k[2], k[3], k[4], k[5], k[6], k[7], MLIIS_ZS);
  }
  // stem (efficientnet_model.py:410-412); its BN+swish is fused into block 0's depthwise loader
  stem_fwd(images, index, r.T(p.w_stem), r.W(p.S0.off), B, p.image_size, p.image_size, p.Hs, p.Ws, p.stem_pad_t,
           p.stem_pad_l, st);
  if (training) r.bn_train(p.bn_stem, r.W(p.S0.off), 32, B * p.Hs * p.Ws, false);

  const float* X = nullptr;   // materialised block input (output of the previous block)
  int Xld = 0;
  for (size_t i = 0; i < p.blocks.size(); ++i) {
    const BlockPlan& b = p.blocks[i];
    const int Mi = B * b.Hin * b.Win, Mo = B * b.Hout * b.Wout, HWo = b.Hout * b.Wout;
    const float *dw_in, *dw_a, *dw_b;
    if (b.expand) {
      Dense{r}.fwd(X, Xld, 0, b.Hin, b.Win, b.cin, 1, r.T(b.w_expand), nullptr, r.W(b.E.off), b.ce, b.ce, Mi,
                   b.Hin * b.Win);
      if (training) r.bn_train(b.bn0, r.W(b.E.off), b.ce, Mi, false);
      dw_in = r.W(b.E.off); dw_a = r.bn_a(b.bn0); dw_b = r.bn_b(b.bn0);
    } else {
      dw_in = r.W(p.S0.off); dw_a = r.bn_a(p.bn_stem); dw_b = r.bn_b(p.bn_stem);
    }
    dw_fwd(dw_in, dw_a, dw_b, r.T(b.w_dw), r.W(b.D.off), B, b.Hin, b.Win, b.ce, b.k, b.stride, b.Hout, b.Wout, b.pad_t,
           b.pad_l, st);
    if (training) r.bn_train(b.bn1, r.W(b.D.off), b.ce, Mo, false);
    // squeeze-excite (efficientnet_model.py:238-251)
    se_pool(r.W(b.D.off), b.ce, r.bn_a(b.bn1), r.bn_b(b.bn1), B, HWo, b.ce, r.W(p.partials), st);
    se_fc_fwd(r.W(p.partials), rc_num_img_chunks(HWo, b.ce), B, HWo, b.ce, b.cr, r.T(b.w_se1), r.T(b.b_se1),
              r.T(b.w_se2), r.T(b.b_se2), r.W(b.pool), r.W(b.hidpre), r.W(b.gate), st);
    // project conv consumes swish(BN1(dw)) * gate, recomputed in the A-operand loader
    Dense{r}.fwd_pro(r.W(b.D.off), b.ce, b.ce, r.T(b.w_proj), r.W(b.P.off), b.cout, b.cout, Mo, HWo, r.bn_a(b.bn1),
                     r.bn_b(b.bn1), r.W(b.gate));
    if (training) r.bn_train(b.bn2, r.W(b.P.off), b.cout, Mo, false);
    const float* dcs = (training && b.dc_idx >= 0) ? r.W(p.dcs + (int64_t)b.dc_idx * p.maxB) : nullptr;
    block_out(r.W(b.P.off), b.cout, r.bn_a(b.bn2), r.bn_b(b.bn2), dcs, b.skip ? X : nullptr, Xld, r.W(b.Y.off),
              b.cout, Mo, b.cout, HWo, st);
    X = r.W(b.Y.off);
    Xld = b.cout;
  }

  // decoder (efficientlab.py:153-231)
  const float* deep = X;
  int deep_ld = Xld;
  for (const RsdPlan& d : p.rsds) {
    const int HW = d.h * d.w, M = B * HW, D = d.D;
    float* cat = r.W(d.cat.off);
    if (d.identity_up)
      add3(cat, d.catC, deep, deep_ld, nullptr, 0, nullptr, 0, M, D, HW, st);
    else
      bilinear_fwd(deep, deep_ld, cat, d.catC, B, d.hin, d.win, d.h, d.w, D, r.c->tabs[d.tab].rt(),
                   r.c->tabs[d.tab].rt(), st);
    const BlockPlan& sb = p.blocks[d.skip_block];
    add3(cat + D, d.catC, r.W(sb.Y.off), sb.cout, nullptr, 0, nullptr, 0, M, d.skipC, HW, st);
    float* pyr = r.W(d.pyr.off);
    // branch_0: 1x1 (+bias) -> swish -> BN
    const Dense dense{r, true};
    dense.fwd(cat, d.catC, 0, d.h, d.w, d.catC, 1, r.T(d.w0), r.T(d.b0), r.W(d.c0.off), D, D, M, HW);
    if (training) r.bn_train(d.bn[0], r.W(d.c0.off), D, M, true);
    dec_bn_apply(r.W(d.c0.off), D, r.bn_a(d.bn[0]), r.bn_b(d.bn[0]), nullptr, 0, pyr, d.pyrC, M, D, st);
    // branch_1: 3x3 dilation 2
    dense.fwd(cat, d.catC, 1, d.h, d.w, d.catC, 2, r.T(d.w1), r.T(d.b1), r.W(d.c1.off), D, D, M, HW);
    if (training) r.bn_train(d.bn[1], r.W(d.c1.off), D, M, true);
    dec_bn_apply(r.W(d.c1.off), D, r.bn_a(d.bn[1]), r.bn_b(d.bn[1]), nullptr, 0, pyr + D, d.pyrC, M, D, st);
    // branch_2: image-level mean (efficientlab.py:192-197).  Tensor-core modes FOLD its tile into conv2d_2 as a
    // per-image, per-border-class bias (k_pool.cu); the fp32 reference mode materialises it like the reference does.
    img_colsum(cat, d.catC, B, HW, d.catC, 1.f / (float)HW, r.W(p.partials), r.W(d.pooled), d.catC, st);
    bool folded = false;
    if (dense.fold() && d.h >= 2 && d.w >= 2) {
      pool_bias9(r.W(d.pooled), d.catC, r.T(d.w2), d.pyrC, 2 * D, d.catC, D, B, r.W(d.bias9), st);
      folded = dense.fwd_fold(pyr, d.pyrC, d.h, d.w, 2 * D, d.pyrC, r.T(d.w2), r.T(d.b2), r.W(d.bias9), r.W(d.c2.off), D,
                              D, M);
    }
    if (!folded) {
      bcast_rows(r.W(d.pooled), d.catC, pyr + 2 * D, d.pyrC, B, HW, d.catC, st);
      // 3x3 over the pyramid, + residual
      dense.fwd(pyr, d.pyrC, 1, d.h, d.w, d.pyrC, 1, r.T(d.w2), r.T(d.b2), r.W(d.c2.off), D, D, M, HW);
    }
    r.sl->folded[&d - &p.rsds[0]] = folded;
    if (training) r.bn_train(d.bn[2], r.W(d.c2.off), D, M, true);
    dec_bn_apply(r.W(d.c2.off), D, r.bn_a(d.bn[2]), r.bn_b(d.bn[2]), cat, d.catC, r.W(d.out.off), D, M, D, st);
    deep = r.W(d.out.off);
    deep_ld = D;
  }
  // head: dropout -> 1x1 -> (bilinear + softmax are fused into the loss / predict kernels)
  const float rate = r.c->cfg.final_dropout_rate;
  const float* mask = nullptr;
  if (training && rate > 0.f) {
    if (drop_mask) mask = drop_mask;
    else {
      fill_dropout_mask(r.W(p.dropmask), (int64_t)B * p.hl * p.wl * p.D, rate, seed, seed_dev, st);
      mask = r.W(p.dropmask);
    }
  }
  r.sl->fwd_drop_mask = mask;
  if (p.n_out == 2) {
    head_fwd(deep, deep_ld, r.T(p.w_head), r.T(p.b_head), mask, 1.f / (1.f - rate), r.W(p.z_lo), B * p.hl * p.wl, p.D,
             st);
  } else {
    // multi-class head (joint training): a dense 1x1 layer D -> Cp on the tensor-core path, weights zero-padded
    // from n_out to Cp = 4*ceil(n_out/4) columns so that every operand row is 16-byte aligned
    const int M = B * p.hl * p.wl;
    mc_pad_head(r.T(p.w_head), r.T(p.b_head), r.W(p.mc_wp), r.W(p.mc_bp), p.D, p.n_out, p.Cp, st);
    const float* xin = deep;
    if (mask) {
      mc_mul_mask(deep, mask, 1.f / (1.f - rate), r.W(p.mc_xdrop), (int64_t)M * p.D, st);
      xin = r.W(p.mc_xdrop);
    }
    Dense{r, true}.fwd(xin, p.D, 0, p.hl, p.wl, p.D, 1, r.W(p.mc_wp), r.W(p.mc_bp), r.W(p.z_lo), p.Cp, p.Cp, M,
                       p.hl * p.wl);
  }
  r.sl->fwd_batch = B;
  r.sl->fwd_training = training;
  r.sl->fwd_images = images;
  r.sl->fwd_index = index;
}

// ------------------------------------------------------------------------------------------------
// loss + backward
// ------------------------------------------------------------------------------------------------
McLossArgs mc_args(const Run& r, const float* mask, const int32_t* index) {
  const Plan& p = r.p;
  const Tab& tf = r.c->tabs[p.tab_final];
  McLossArgs a{};
  a.z_lo = r.W(p.z_lo); a.ldz = p.Cp; a.mask = mask; a.cls = r.sl->class_ids; a.index = index;
  a.B = r.B; a.h = p.hl; a.w = p.wl; a.H = p.image_size; a.W = p.image_size; a.C = p.n_out;
  a.ty = tf.rt(); a.tx = tf.rt();
  a.dice = (r.c->cfg.loss_flags & MLIIS_LOSS_DICE) ? 1 : 0;
  a.label_smoothing = r.c->cfg.label_smoothing;
  a.lse = r.W(p.mc_lse); a.pt = r.W(p.p1); a.partials = r.W(p.partials); a.coef = r.W(p.loss_coef);
  a.dz_lo = r.W(p.dz_lo); a.lddz = p.Cp; a.cellgrad = r.W(p.mc_cellgrad);
  a.theta = r.theta; a.n_l2 = p.n_l2;
  a.l2_coef = (r.c->cfg.loss_flags & MLIIS_LOSS_L2) ? 0.0005f : 0.f;
  return a;
}

// multi-class loss (sparse labels) + head backward; leaves d loss / d features in g_out like the binary head
void run_backward_multiclass_head(const Run& r, const float* mask, const int32_t* index, float* loss_out) {
  const Plan& p = r.p;
  cudaStream_t st = r.st;
  const int M = r.B * p.hl * p.wl, HW = p.hl * p.wl;
  McLossArgs a = mc_args(r, mask, index);
  a.loss_out = loss_out;
  mc_loss_fwd_bwd(a, st);
  const RsdPlan& last = p.rsds.back();
  const float rate = r.c->cfg.final_dropout_rate;
  const float* dmask = r.sl->fwd_drop_mask;
  const float* xin = dmask ? r.W(p.mc_xdrop) : r.W(last.out.off);
  const Dense dense{r, true};
  // dW^T via the swapped-role wgrad (Cp > 256 output channels), db = column sums of dz
  dense.wgrad(xin, p.D, 0, p.hl, p.wl, p.D, 1, r.W(p.dz_lo), p.Cp, p.Cp, r.W(p.mc_gwp), r.W(p.mc_gbp), M, HW,
              nullptr, nullptr, nullptr, true);
  dense.dgrad(r.W(p.dz_lo), p.Cp, 0, p.hl, p.wl, p.D, 1, r.W(p.mc_wp), r.W(p.g_out), p.D, p.Cp, M, HW, 0);
  if (dmask) mc_mul_mask(r.W(p.g_out), dmask, 1.f / (1.f - rate), r.W(p.g_out), (int64_t)M * p.D, st);
  mc_unpad_grad(r.W(p.mc_gwp), r.W(p.mc_gbp), r.G(p.w_head), r.G(p.b_head), p.D, p.n_out, p.Cp, st);
}

void run_backward(const Run& r, const float* labels, const int32_t* index, float* loss_out) {
  const Plan& p = r.p;
  const int B = r.B;
  cudaStream_t st = r.st;
  const Tab& tf = r.c->tabs[p.tab_final];
  for (int z = 0; z < MLIIS_NZ; ++z) {   // padding holes of the flat gradient buffer must read as zero
    Slot& sz = r.sl[z];
    if (!sz.grads_zeroed) {
      cudaMemsetAsync(r.G(0) + (size_t)z * MLIIS_ZS, 0, p.n_theta * sizeof(float), st);
      sz.grads_zeroed = true;
    }
  }
  if (p.n_out != 2) {
    run_backward_multiclass_head(r, labels, index, loss_out);
  } else {
  LossArgs la{};
  la.z_lo = r.W(p.z_lo); la.labels = labels; la.index = index;
  la.B = B; la.h = p.hl; la.w = p.wl; la.H = p.image_size; la.W = p.image_size;
  la.ty = tf.rt(); la.tx = tf.rt();
  la.dice = (r.c->cfg.loss_flags & MLIIS_LOSS_DICE) ? 1 : 0;
  la.label_smoothing = r.c->cfg.label_smoothing;
  la.p1 = r.W(p.p1); la.partials = r.W(p.partials); la.coef = r.W(p.loss_coef); la.dz_hi = r.W(p.dz_hi);
  la.loss_out = loss_out;
  la.theta = r.theta; la.n_l2 = p.n_l2;
  la.l2_coef = (r.c->cfg.loss_flags & MLIIS_LOSS_L2) ? 0.0005f : 0.f;
  loss_fwd_bwd(la, st);
  bilinear_bwd(r.W(p.dz_hi), 2, r.W(p.dz_lo), 2, B, p.hl, p.wl, p.image_size, p.image_size, 2, tf.rt(), tf.rt(), st);

  // head
  const RsdPlan& last = p.rsds.back();
  const float rate = r.c->cfg.final_dropout_rate;
  head_bwd(r.W(last.out.off), p.D, r.T(p.w_head), r.sl->fwd_drop_mask, 1.f / (1.f - rate), r.W(p.dz_lo), r.W(p.g_out),
           p.D, r.W(p.partials), r.G(p.w_head), r.G(p.b_head), B * p.hl * p.wl, p.D, st);
  }

  // decoder, reverse order
  float* gOut = r.W(p.g_out);
  for (int di = (int)p.rsds.size() - 1; di >= 0; --di) {
    const RsdPlan& d = p.rsds[di];
    const int HW = d.h * d.w, M = B * HW, D = d.D;
    const float* cat = r.W(d.cat.off);
    const float* pyr = r.W(d.pyr.off);
    auto dec_bn_bwd = [&](const BnRef& bn, const float* x, const float* g, int ldg, float* dx) {
      BnBwdArgs a{};
      a.x = x; a.ldx = D; a.g = g; a.ldg = ldg; a.dx = dx; a.lddx = D; a.M = M; a.C = D; a.HW = HW;
      a.mean = r.bn_mean(bn); a.rstd = r.bn_rstd(bn); a.a = r.bn_a(bn); a.b = r.bn_b(bn); a.gamma = r.T(bn.gamma);
      a.partials = r.W(p.partials); a.k = r.W(p.bn_k); a.ticket = r.ticket(); a.dgamma = r.G(bn.gamma); a.dbeta = r.G(bn.beta);
      bn_bwd(BN_DEC, a, st);
    };
    // out = BN2(swish(c2)) + up
    dec_bn_bwd(d.bn[2], r.W(d.c2.off), gOut, D, r.W(p.g_c));
    const Dense dense{r, true};
    const float* gpyr = r.W(p.g_pyr);
    if (r.sl->folded[di]) {
      // conv2d_2 over the 2D real channels; the pooled channels' weight / input gradients come from the region sums
      // of dL/d(conv2d_2) (k_pool.cu)
      dense.wgrad(pyr, d.pyrC, 1, d.h, d.w, 2 * D, 1, r.W(p.g_c), D, D, r.G(d.w2), r.G(d.b2), M, HW, nullptr, nullptr,
                  nullptr, false, d.pyrC * D);
      region_sums(r.W(p.g_c), D, B, d.h, d.w, 1, D, r.W(p.partials), r.W(d.S9), st);
      pool_wgrad(r.W(d.pooled), d.catC, r.W(d.S9), B, d.pyrC, 2 * D, d.catC, D, r.G(d.w2), st);
      dense.dgrad(r.W(p.g_c), D, 1, d.h, d.w, 2 * D, 1, r.T(d.w2), r.W(p.g_pyr), d.pyrC, D, M, HW, 0, d.pyrC);
      pool_dgrad(r.W(d.S9), r.T(d.w2), d.pyrC, 2 * D, d.catC, D, B, 1.f / (float)HW, r.W(d.dpooled), d.catC, st);
    } else {
      dense.wgrad(pyr, d.pyrC, 1, d.h, d.w, d.pyrC, 1, r.W(p.g_c), D, D, r.G(d.w2), r.G(d.b2), M, HW);
      dense.dgrad(r.W(p.g_c), D, 1, d.h, d.w, d.pyrC, 1, r.T(d.w2), r.W(p.g_pyr), d.pyrC, D, M, HW, 0);
      img_colsum(gpyr + 2 * D, d.pyrC, B, HW, d.catC, 1.f / (float)HW, r.W(p.partials), r.W(d.dpooled), d.catC, st);
    }
    dec_bn_bwd(d.bn[0], r.W(d.c0.off), gpyr, d.pyrC, r.W(p.g_c0));
    dec_bn_bwd(d.bn[1], r.W(d.c1.off), gpyr + D, d.pyrC, r.W(p.g_c1));
    // branch_0 1x1
    dense.wgrad(cat, d.catC, 0, d.h, d.w, d.catC, 1, r.W(p.g_c0), D, D, r.G(d.w0), r.G(d.b0), M, HW);
    dense.dgrad(r.W(p.g_c0), D, 0, d.h, d.w, d.catC, 1, r.T(d.w0), r.W(p.g_cat), d.catC, D, M, HW, 0);
    // branch_1 3x3 dil 2 (accumulates into g_cat)
    dense.wgrad(cat, d.catC, 1, d.h, d.w, d.catC, 2, r.W(p.g_c1), D, D, r.G(d.w1), r.G(d.b1), M, HW);
    dense.dgrad(r.W(p.g_c1), D, 1, d.h, d.w, d.catC, 2, r.T(d.w1), r.W(p.g_cat), d.catC, D, M, HW, 1);
    // d_up = gOut + g_cat[:, :D] + dpooled[:, :D]      d_skip = g_cat[:, D:] + dpooled[:, D:]
    const float* gcat = r.W(p.g_cat);
    const float* dpl = r.W(d.dpooled);
    add3(r.W(p.g_up), D, gOut, D, gcat, d.catC, dpl, d.catC, M, D, HW, st);
    const BlockPlan& sb = p.blocks[d.skip_block];
    float* extra = r.W(sb.extra_grad);
    if (d.identity_up && di == 0 && d.skip_block == (int)p.blocks.size() - 1) {
      // deep input and skip are the same tensor (reduction_4): both gradients land on it
      add3(r.W(p.g_skip), d.skipC, gcat + D, d.catC, nullptr, 0, dpl + D, d.catC, M, d.skipC, HW, st);
      add3(extra, D, r.W(p.g_up), D, r.W(p.g_skip), d.skipC, nullptr, 0, M, D, HW, st);
    } else {
      add3(extra, d.skipC, gcat + D, d.catC, nullptr, 0, dpl + D, d.catC, M, d.skipC, HW, st);
      float* gdeep = di == 0 ? r.W(p.blocks.back().extra_grad) : r.W(p.g_deep);
      if (d.identity_up)
        add3(gdeep, D, r.W(p.g_up), D, nullptr, 0, nullptr, 0, M, D, HW, st);
      else
        bilinear_bwd(r.W(p.g_up), D, gdeep, D, B, d.hin, d.win, d.h, d.w, D, r.c->tabs[d.tab].rt(),
                     r.c->tabs[d.tab].rt(), st);
      gOut = gdeep;
    }
  }

  // backbone, reverse order.  gY[cur] holds dL/dY of the current block.
  int cur = 0;
  const int nb = (int)p.blocks.size();
  for (int i = nb - 1; i >= 0; --i) {
    const BlockPlan& b = p.blocks[i];
    const int Mi = B * b.Hin * b.Win, Mo = B * b.Hout * b.Wout, HWo = b.Hout * b.Wout, HWi = b.Hin * b.Win;
    float* gY = r.W(p.gY[cur]);
    if (i == nb - 1) {
      gY = r.W(b.extra_grad);
    } else if (b.extra_grad >= 0) {
      add3(gY, b.cout, gY, b.cout, r.W(b.extra_grad), b.cout, nullptr, 0, Mo, b.cout, HWo, st);
    }
    const float* X = i > 0 ? r.W(p.blocks[i - 1].Y.off) : nullptr;
    // Y = BN2(P) * dcs + X
    {
      BnBwdArgs a{};
      a.x = r.W(b.P.off); a.ldx = b.cout; a.g = gY; a.ldg = b.cout; a.dx = r.W(p.gP); a.lddx = b.cout;
      a.M = Mo; a.C = b.cout; a.HW = HWo;
      a.mean = r.bn_mean(b.bn2); a.rstd = r.bn_rstd(b.bn2); a.a = r.bn_a(b.bn2); a.b = r.bn_b(b.bn2);
      a.gamma = r.T(b.bn2.gamma);
      a.dcs = b.dc_idx >= 0 ? r.W(p.dcs + (int64_t)b.dc_idx * p.maxB) : nullptr;
      a.partials = r.W(p.partials); a.k = r.W(p.bn_k); a.ticket = r.ticket(); a.dgamma = r.G(b.bn2.gamma); a.dbeta = r.G(b.bn2.beta);
      bn_bwd(BN_PLAIN, a, st);
    }
    // project conv: wgrad on swish(BN1(D))*gate (recomputed), dgrad into gD
    Dense{r}.wgrad(r.W(b.D.off), b.ce, 0, b.Hout, b.Wout, b.ce, 1, r.W(p.gP), b.cout, b.cout, r.G(b.w_proj), nullptr, Mo,
                   HWo, r.bn_a(b.bn1), r.bn_b(b.bn1), r.W(b.gate));
    Dense{r}.dgrad(r.W(p.gP), b.cout, 0, b.Hout, b.Wout, b.ce, 1, r.T(b.w_proj), r.W(p.gD), b.ce, b.cout, Mo, HWo, 0);
    // squeeze-excite backward
    se_bwd_reduce(r.W(b.D.off), b.ce, r.W(p.gD), b.ce, r.bn_a(b.bn1), r.bn_b(b.bn1), B, HWo, b.ce, r.W(p.partials), st);
    se_fc_bwd(r.W(p.partials), rc_num_img_chunks(HWo, b.ce), B, HWo, b.ce, b.cr, r.T(b.w_se1), r.T(b.w_se2),
              r.W(b.pool), r.W(b.hidpre), r.W(b.gate), r.G(b.w_se1), r.G(b.b_se1), r.G(b.w_se2), r.G(b.b_se2),
              r.W(b.dpool), st);
    {
      BnBwdArgs a{};
      a.x = r.W(b.D.off); a.ldx = b.ce; a.g = r.W(p.gD); a.ldg = b.ce; a.dx = r.W(p.gD); a.lddx = b.ce;
      a.M = Mo; a.C = b.ce; a.HW = HWo;
      a.mean = r.bn_mean(b.bn1); a.rstd = r.bn_rstd(b.bn1); a.a = r.bn_a(b.bn1); a.b = r.bn_b(b.bn1);
      a.gamma = r.T(b.bn1.gamma); a.gate = r.W(b.gate); a.dpool = r.W(b.dpool);
      a.partials = r.W(p.partials); a.k = r.W(p.bn_k); a.ticket = r.ticket(); a.dgamma = r.G(b.bn1.gamma); a.dbeta = r.G(b.bn1.beta);
      bn_bwd(BN_SWISH_SE, a, st);
    }
    // depthwise
    const float *dw_in, *dw_a, *dw_b;
    const BnRef& bnin = b.expand ? b.bn0 : p.bn_stem;
    dw_in = b.expand ? r.W(b.E.off) : r.W(p.S0.off);
    dw_a = r.bn_a(bnin); dw_b = r.bn_b(bnin);
    dw_bwd_weight(dw_in, dw_a, dw_b, r.W(p.gD), r.W(p.partials), r.G(b.w_dw), B, b.Hin, b.Win, b.ce, b.k, b.stride,
                  b.Hout, b.Wout, b.pad_t, b.pad_l, st);
    dw_bwd_data(r.W(p.gD), r.T(b.w_dw), r.W(p.gE), B, b.Hin, b.Win, b.ce, b.k, b.stride, b.Hout, b.Wout, b.pad_t,
                b.pad_l, st);
    {
      BnBwdArgs a{};
      a.x = dw_in; a.ldx = b.ce; a.g = r.W(p.gE); a.ldg = b.ce; a.dx = r.W(p.gE); a.lddx = b.ce;
      a.M = Mi; a.C = b.ce; a.HW = HWi;
      a.mean = r.bn_mean(bnin); a.rstd = r.bn_rstd(bnin); a.a = dw_a; a.b = dw_b; a.gamma = r.T(bnin.gamma);
      a.partials = r.W(p.partials); a.k = r.W(p.bn_k); a.ticket = r.ticket(); a.dgamma = r.G(bnin.gamma); a.dbeta = r.G(bnin.beta);
      bn_bwd(BN_SWISH, a, st);
    }
    if (b.expand) {
      Dense{r}.wgrad(X, b.cin, 0, b.Hin, b.Win, b.cin, 1, r.W(p.gE), b.ce, b.ce, r.G(b.w_expand), nullptr, Mi, HWi,
                     nullptr, nullptr, nullptr, /*swap=*/true);
      // dX = dE * We^T (+ dY through the identity skip: same shape, accumulate in place)
      float* gX = b.skip ? gY : r.W(p.gY[cur ^ 1]);
      Dense{r}.dgrad(r.W(p.gE), b.ce, 0, b.Hin, b.Win, b.cin, 1, r.T(b.w_expand), gX, b.cin, b.ce, Mi, HWi,
                     b.skip ? 1 : 0);
      if (b.skip && gY != r.W(p.gY[cur])) {
        // gradient lived in the decoder's extra buffer: move it into the ping-pong chain
        add3(r.W(p.gY[cur]), b.cin, gX, b.cin, nullptr, 0, nullptr, 0, Mi, b.cin, HWi, st);
      } else if (!b.skip) {
        cur ^= 1;
      }
    } else {
      // block 0: gE is dL/d(stem conv output) after the stem BN backward above
      stem_wgrad(r.sl->fwd_images, r.sl->fwd_index, r.W(p.gE), r.W(p.partials), r.G(p.w_stem), B, p.image_size,
                 p.image_size, p.Hs, p.Ws, p.stem_pad_t, p.stem_pad_l, st);
    }
  }
}

int validate(mliis_ctx* ctx, int slot, int batch) {
  if (!ctx) return fail(MLIIS_ERR_ARG, "null ctx");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx (no sm_100 device): there is no CPU fallback");
  if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(MLIIS_ERR_ARG, "slot %d out of range", slot);
  if (!ctx->slots[slot].state || !ctx->slots[slot].ws) return fail(MLIIS_ERR_STATE, "slot %d not bound", slot);
  if (batch < 1 || batch > ctx->plan.maxB) return fail(MLIIS_ERR_ARG, "batch %d not in [1, %d]", batch, ctx->plan.maxB);
  return MLIIS_OK;
}

void set_lr(const Run& r, float lr) {
  MLIIS_COUNT(), set_scalar_kernel<<<dim3(1, 1, MLIIS_NZ), 1, 0, r.st>>>(r.W(r.p.lr_dev), lr, MLIIS_ZS);
}

void run_optimizer(const Run& r, const float* lr_dev) {
  const Plan& p = r.p;
  const float l2 = (r.c->cfg.loss_flags & MLIIS_LOSS_L2) ? 0.0005f : 0.f;
  adam_step(r.theta, r.adam_v, r.G(0), p.n_theta, p.n_l2, lr_dev, r.powers, l2, r.c->cfg.optimizer == MLIIS_OPT_SGD,
            r.st);
}

}  // namespace

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

const char* mliis_last_error(void) { return g_err.c_str(); }
const char* mliis_version(void) { return "mliis_b200 0.1 (sm_100a)"; }

int mliis_ctx_create(const mliis_config* cfg, int device, mliis_ctx** out) {
  if (!cfg || !out) return fail(MLIIS_ERR_ARG, "null argument");
  mliis_ctx* c = new mliis_ctx();
  c->cfg = *cfg;
  try {
    if (cfg->n_classes > 1 && cfg->max_batch > 256)
      throw std::invalid_argument("multi-class head: max_batch must be <= 256");
    c->plan.build(cfg->image_size, cfg->max_batch, cfg->rsd, cfg->final_dropout_rate,
                  cfg->n_classes > 1 ? cfg->n_classes + 1 : 2);
  } catch (const std::exception& e) {
    delete c;
    return fail(MLIIS_ERR_ARG, "%s", e.what());
  }
  if (cfg->n_slots < 1 || cfg->n_slots > 1024) { delete c; return fail(MLIIS_ERR_ARG, "n_slots must be in [1,1024]"); }
  if (cfg->final_dropout_rate < 0.f || cfg->final_dropout_rate >= 1.f) { delete c; return fail(MLIIS_ERR_ARG, "bad dropout rate"); }
  if (c->plan.n_dc > 8) { delete c; return fail(MLIIS_ERR_ARG, "too many drop-connect blocks"); }
  c->slots.resize(cfg->n_slots);
  for (int i = 0; i < 16; ++i) c->keep[i] = 1.f;
  for (const BlockPlan& b : c->plan.blocks)
    if (b.dc_idx >= 0) c->keep[b.dc_idx] = 1.f - b.dc_rate;
  c->device = device;
  if (device >= 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
      delete c;
      return fail(MLIIS_ERR_DEVICE, "no CUDA device %d: %s (no CPU fallback exists)", device,
                  cudaGetErrorString(cudaGetLastError()));
    }
    if (prop.major != 10) {
      delete c;
      return fail(MLIIS_ERR_DEVICE, "device %d is sm_%d%d, this library is sm_100a only (no fallback)", device,
                  prop.major, prop.minor);
    }
    cudaSetDevice(device);
    const Plan& p = c->plan;
    std::vector<int32_t> gi(p.n_bn_ch), bi(p.n_bn_ch);
    auto fill = [&](const BnRef& r) {
      if (r.idx < 0) return;
      for (int ch = 0; ch < r.C; ++ch) { gi[r.off + ch] = (int32_t)(r.gamma + ch); bi[r.off + ch] = (int32_t)(r.beta + ch); }
    };
    fill(p.bn_stem);
    for (const BlockPlan& b : p.blocks) { fill(b.bn0); fill(b.bn1); fill(b.bn2); }
    for (const RsdPlan& d : p.rsds) for (int j = 0; j < 3; ++j) fill(d.bn[j]);
    c->d_gamma_idx = upload(c, gi);
    c->d_beta_idx = upload(c, bi);
    if (c->cfg.gemm_mode != MLIIS_GEMM_FP32 && !p.prep_jobs.empty()) {
      std::vector<TcPrepJob> jobs;
      for (const Plan::PrepJob& j : p.prep_jobs)
        jobs.push_back({(long long)j.w_off, (long long)j.dst, j.taps, j.Ci, j.Co, j.dgrad,
                        (c->cfg.gemm_mode == MLIIS_GEMM_TF32 && j.decoder) ? 1 : 3, j.Cs});
      c->d_prep_jobs = upload(c, jobs);
    }
    c->tabs.resize(p.resize_pairs.size());
    bool ok = c->d_gamma_idx && c->d_beta_idx;
    for (size_t i = 0; ok && i < p.resize_pairs.size(); ++i)
      ok = make_tab(c, p.resize_pairs[i].first, p.resize_pairs[i].second, &c->tabs[i]);
    if (!ok) { mliis_ctx_destroy(c); return fail(MLIIS_ERR_CUDA, "device table allocation failed"); }
  }
  *out = c;
  return MLIIS_OK;
}

int mliis_ctx_destroy(mliis_ctx* ctx) {
  if (!ctx) return MLIIS_OK;
  for (Slot& sl : ctx->slots) {
    if (sl.graph_exec) cudaGraphExecDestroy(sl.graph_exec);
    if (sl.graph) cudaGraphDestroy(sl.graph);
  }
  for (void* p : ctx->owned) cudaFree(p);
  if (ctx->nccl_comm) mliis_comm_destroy(ctx);
  delete ctx;
  return MLIIS_OK;
}

int64_t mliis_num_params(const mliis_ctx* c) { return c ? c->plan.n_params : -1; }
int32_t mliis_num_param_tensors(const mliis_ctx* c) { return c ? (int32_t)c->plan.params.size() : -1; }
int32_t mliis_num_bn_layers(const mliis_ctx* c) { return c ? (int32_t)c->plan.bns.size() : -1; }
int32_t mliis_num_bn_channels(const mliis_ctx* c) { return c ? c->plan.n_bn_ch : -1; }
int32_t mliis_num_dc_blocks(const mliis_ctx* c) { return c ? c->plan.n_dc : -1; }

int mliis_param_table(const mliis_ctx* c, mliis_param_info* out, int32_t capacity) {
  if (!c || !out) return fail(MLIIS_ERR_ARG, "null argument");
  if (capacity < (int)c->plan.params.size()) return fail(MLIIS_ERR_ARG, "capacity too small");
  for (size_t i = 0; i < c->plan.params.size(); ++i) {
    const ParamEntry& e = c->plan.params[i];
    out[i].name = e.name.c_str();
    out[i].offset = e.offset;
    out[i].size = e.size;
    out[i].ndim = e.ndim;
    for (int j = 0; j < 4; ++j) out[i].shape[j] = e.shape[j];
    out[i].l2 = e.l2;
  }
  return MLIIS_OK;
}

int mliis_bn_table(const mliis_ctx* c, mliis_bn_info* out, int32_t capacity) {
  if (!c || !out) return fail(MLIIS_ERR_ARG, "null argument");
  if (capacity < (int)c->plan.bns.size()) return fail(MLIIS_ERR_ARG, "capacity too small");
  for (size_t i = 0; i < c->plan.bns.size(); ++i) {
    out[i].scope = c->plan.bns[i].scope.c_str();
    out[i].channels = c->plan.bns[i].C;
    out[i].offset = c->plan.bns[i].off;
    out[i].fused = c->plan.bns[i].fused;
  }
  return MLIIS_OK;
}

int64_t mliis_workspace_bytes(const mliis_ctx* c) { return c ? c->plan.ws_floats * (int64_t)sizeof(float) : -1; }
int64_t mliis_state_floats(const mliis_ctx* c) {
  return c ? 2 * c->plan.n_theta + 2 * (int64_t)c->plan.n_bn_ch + 4 : -1;
}
int64_t mliis_theta_floats(const mliis_ctx* c) { return c ? c->plan.n_theta : -1; }

int mliis_slot_bind(mliis_ctx* ctx, int32_t slot, float* dev_state, void* dev_workspace) {
  if (!ctx) return fail(MLIIS_ERR_ARG, "null ctx");
  if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(MLIIS_ERR_ARG, "slot out of range");
  if (((uintptr_t)dev_state & 15) || ((uintptr_t)dev_workspace & 255)) return fail(MLIIS_ERR_ARG, "misaligned buffers");
  ctx->slots[slot] = Slot();
  ctx->slots[slot].state = dev_state;
  ctx->slots[slot].ws = (float*)dev_workspace;
  if (ctx->device >= 0) {      // the ticket words of the clustered reductions start at zero (they reset themselves afterwards)
    const Plan& p = ctx->plan;
    cudaMemset((float*)dev_workspace + p.partials + p.partials_len - 16, 0, 16 * sizeof(float));
  }
  return MLIIS_OK;
}

int mliis_state_copy(mliis_ctx* ctx, float* dst, const float* src, int32_t what, void* stream) {
  if (!ctx || !dst || !src) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx");
  const Plan& p = ctx->plan;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t o_bn = p.n_theta, o_opt = p.n_theta + 2 * p.n_bn_ch;
  if (what & MLIIS_STATE_TRAINABLES) cudaMemcpyAsync(dst, src, p.n_theta * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (what & MLIIS_STATE_BN)
    cudaMemcpyAsync(dst + o_bn, src + o_bn, 2 * (size_t)p.n_bn_ch * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (what & MLIIS_STATE_OPT)
    cudaMemcpyAsync(dst + o_opt, src + o_opt, ((size_t)p.n_theta + 4) * sizeof(float), cudaMemcpyDeviceToDevice, st);
  return check_cuda("state_copy");
}

int mliis_forward(mliis_ctx* ctx, int32_t slot, const float* images, const int32_t* index, int32_t batch,
                  int32_t training, const float* dc_mask, const float* drop_mask, uint64_t seed, float* logits_out,
                  void* stream) {
  int rc = validate(ctx, slot, batch);
  if (rc) return rc;
  if (!images) return fail(MLIIS_ERR_ARG, "null images");
  Run r(ctx, slot, batch, (cudaStream_t)stream);
  run_forward(r, images, index, training != 0, dc_mask, drop_mask, seed);
  if (logits_out && ctx->plan.n_out != 2) {
    // multi-class head: the full-resolution logits are never materialised; return the LOW-resolution head output
    // [B, h, w, n_out] (the loss / predict kernels upsample on the fly)
    const Plan& p = ctx->plan;
    cudaMemcpy2DAsync(logits_out, (size_t)p.n_out * sizeof(float), r.W(p.z_lo), (size_t)p.Cp * sizeof(float),
                      (size_t)p.n_out * sizeof(float), (size_t)batch * p.hl * p.wl, cudaMemcpyDeviceToDevice, r.st);
  } else if (logits_out) {
    const Plan& p = ctx->plan;
    const Tab& tf = ctx->tabs[p.tab_final];
    predict_mask_iou(r.W(p.z_lo), nullptr, nullptr, batch, p.hl, p.wl, p.image_size, p.image_size, tf.rt(), tf.rt(),
                     nullptr, logits_out, nullptr, nullptr, r.st);
  }
  return check_cuda("forward");
}

int mliis_loss_backward(mliis_ctx* ctx, int32_t slot, const float* labels, const int32_t* index, int32_t batch,
                        float* grads_out, float* loss_out, void* stream) {
  int rc = validate(ctx, slot, batch);
  if (rc) return rc;
  Slot& sl = ctx->slots[slot];
  if (!sl.fwd_training || sl.fwd_batch != batch) return fail(MLIIS_ERR_STATE, "loss_backward needs a training forward of the same batch");
  if (!labels) return fail(MLIIS_ERR_ARG, "null labels");
  if (ctx->plan.n_out != 2 && !sl.class_ids)
    return fail(MLIIS_ERR_STATE, "multi-class head: call mliis_set_class_ids first (labels = masks [n,H,W])");
  Run r(ctx, slot, batch, (cudaStream_t)stream);
  run_backward(r, labels, index, loss_out);
  if (grads_out)
    cudaMemcpyAsync(grads_out, r.G(0), ctx->plan.n_theta * sizeof(float), cudaMemcpyDeviceToDevice, r.st);
  return check_cuda("loss_backward");
}

int mliis_set_grads(mliis_ctx* ctx, int32_t slot, const float* grads, void* stream) {
  int rc = validate(ctx, slot, 1);
  if (rc) return rc;
  if (!grads) return fail(MLIIS_ERR_ARG, "null grads");
  Run r(ctx, slot, 1, (cudaStream_t)stream);
  cudaMemcpyAsync(r.G(0), grads, ctx->plan.n_theta * sizeof(float), cudaMemcpyDeviceToDevice, r.st);
  return check_cuda("set_grads");
}

int mliis_set_class_ids(mliis_ctx* ctx, int32_t slot, const int32_t* dev_class_ids) {
  int rc = validate(ctx, slot, 1);
  if (rc) return rc;
  if (ctx->plan.n_out == 2) return fail(MLIIS_ERR_STATE, "class ids are only used by the multi-class head (n_classes > 1)");
  ctx->slots[slot].class_ids = dev_class_ids;
  return MLIIS_OK;
}

int mliis_predict_classes(mliis_ctx* ctx, int32_t slot, const float* images, const float* masks, const int32_t* index,
                          int32_t batch, int32_t* class_map_out, uint32_t* inter_out, uint32_t* union_out,
                          void* stream) {
  int rc = validate(ctx, slot, batch);
  if (rc) return rc;
  if (ctx->plan.n_out == 2) return fail(MLIIS_ERR_STATE, "mliis_predict_classes needs the multi-class head (n_classes > 1)");
  if (!images) return fail(MLIIS_ERR_ARG, "null images");
  if ((inter_out || union_out) && !(inter_out && union_out && masks && ctx->slots[slot].class_ids))
    return fail(MLIIS_ERR_ARG, "IoU counts need masks, class ids (mliis_set_class_ids), inter_out and union_out");
  Run r(ctx, slot, batch, (cudaStream_t)stream);
  run_forward(r, images, index, false, nullptr, nullptr, 0);
  McLossArgs a = mc_args(r, masks, index);
  mc_predict(a, class_map_out, inter_out, union_out, r.st);
  return check_cuda("predict_classes");
}

int mliis_optimizer_step(mliis_ctx* ctx, int32_t slot, float lr, float pre_decay_rate, void* stream) {
  int rc = validate(ctx, slot, 1);
  if (rc) return rc;
  (void)pre_decay_rate;
  Run r(ctx, slot, 1, (cudaStream_t)stream);
  set_lr(r, lr);
  run_optimizer(r, r.W(ctx->plan.lr_dev));
  return check_cuda("optimizer_step");
}

static int group_validate(mliis_ctx* ctx, int32_t slot, int32_t n_group, int64_t stride_bytes);

// After mliis_kernel_group(n, stride) the step runs for the n slots slot .. slot + n - 1 in LOCKSTEP (one launch per
// kernel, the slot as a grid dimension): every dev_* pointer of the args is the first slot's, slot k uses pointer +
// k * stride; lr, pre_decay_rate and the host seed are shared.  This is how the slot-parallel meta-TRAINING path
// batches the tasks of a meta-batch (runner.TrainSlots); bit-identical to n single-slot calls.
int mliis_train_step(mliis_ctx* ctx, int32_t slot, const mliis_step_args* a, void* stream) {
  if (!a) return fail(MLIIS_ERR_ARG, "null args");
  int rc = validate(ctx, slot, a->batch);
  if (rc) return rc;
  if (!a->dev_images || !a->dev_labels) return fail(MLIIS_ERR_ARG, "null images/labels");
  if (ctx->plan.n_out != 2 && !ctx->slots[slot].class_ids)
    return fail(MLIIS_ERR_STATE, "multi-class head: call mliis_set_class_ids first (labels = masks [n,H,W])");
  const int ng = t_kernel_group.nz > 1 ? t_kernel_group.nz : 1;
  if (ng > 1 && ctx->plan.n_out != 2) return fail(MLIIS_ERR_STATE, "task-batched steps run on the binary head");
  rc = group_validate(ctx, slot, ng, t_kernel_group.zs * 4);
  if (rc) return rc;
  ZScope zscope(ng, ng > 1 ? t_kernel_group.zs : 0);
  ctx->group_fallback = false;
  Run r(ctx, slot, a->batch, (cudaStream_t)stream);
  // reptile.py:112-113 pre_step_op: var *= rate, before the step's forward pass
  if (a->pre_decay_rate != 1.f && a->pre_decay_rate != 0.f) scale_buffer(r.theta, ctx->plan.n_theta, a->pre_decay_rate, r.st);
  run_forward(r, a->dev_images, a->dev_index, true, a->dev_dc_mask, a->dev_drop_mask, a->seed, a->dev_seed);
  run_backward(r, a->dev_labels, a->dev_index, a->dev_loss_out);
  set_lr(r, a->lr);
  run_optimizer(r, r.W(ctx->plan.lr_dev));
  if (ctx->group_fallback)
    return fail(MLIIS_ERR_STATE, "task-batched execution needs the tensor-core modes (a layer fell back to the "
                                 "single-slot fp32 GEMM kernels)");
  return check_cuda("train_step");
}

int mliis_predict(mliis_ctx* ctx, int32_t slot, const float* images, const float* labels, const int32_t* index,
                  int32_t batch, float* pred_out, float* logits_out, uint32_t* inter_out, uint32_t* union_out,
                  void* stream) {
  int rc = validate(ctx, slot, batch);
  if (rc) return rc;
  if (!images) return fail(MLIIS_ERR_ARG, "null images");
  if (ctx->plan.n_out != 2) return fail(MLIIS_ERR_STATE, "binary head only: use mliis_predict_classes");
  if ((inter_out || union_out) && !(inter_out && union_out && labels))
    return fail(MLIIS_ERR_ARG, "IoU counts need labels, inter_out and union_out");
  Run r(ctx, slot, batch, (cudaStream_t)stream);
  run_forward(r, images, index, false, nullptr, nullptr, 0);
  const Plan& p = ctx->plan;
  const Tab& tf = ctx->tabs[p.tab_final];
  predict_mask_iou(r.W(p.z_lo), labels, index, batch, p.hl, p.wl, p.image_size, p.image_size, tf.rt(), tf.rt(), pred_out,
                   logits_out, inter_out, union_out, r.st);
  return check_cuda("predict");
}

static int task_body(mliis_ctx* ctx, int32_t slot, const mliis_task_args* a, cudaStream_t st) {
  const Plan& p = ctx->plan;
  if (p.n_out != 2) return fail(MLIIS_ERR_STATE, "task adaptation runs on the binary head (n_classes <= 1)");
  const int ng = a->n_group > 1 ? a->n_group : 1;
  ZScope zscope(ng, ng > 1 ? (long long)(a->group_stride_bytes / 4) : 0);   // every launch below serves ng slots
  ctx->group_fallback = false;
  int rc = MLIIS_OK;
  for (int k = 0; k < ng && !rc; ++k)
    rc = mliis_state_copy(ctx, ctx->slots[slot + k].state, a->dev_init_state, MLIIS_STATE_ALL, (void*)st);
  if (rc) return rc;
  for (int t = 0; t < a->n_steps; ++t) {
    Run r(ctx, slot, a->batch, st);
    if (a->pre_decay_rate != 1.f && a->pre_decay_rate != 0.f) scale_buffer(r.theta, p.n_theta, a->pre_decay_rate, st);
    const int32_t* idx = a->dev_batch_index + (size_t)t * a->batch;
    const float* dcm = a->dev_dc_mask ? a->dev_dc_mask + (size_t)t * p.n_dc * a->batch : nullptr;
    run_forward(r, a->dev_images, idx, true, dcm, nullptr, a->seed + (uint64_t)t, a->dev_seed);
    run_backward(r, a->dev_labels, idx, a->dev_loss_out ? a->dev_loss_out + t : nullptr);
    run_optimizer(r, a->dev_lr + t);
  }
  Run r(ctx, slot, a->n_query, st);
  run_forward(r, a->dev_images, a->dev_query_index, false, nullptr, nullptr, 0);
  const Tab& tf = ctx->tabs[p.tab_final];
  predict_mask_iou(r.W(p.z_lo), a->dev_labels, a->dev_query_index, a->n_query, p.hl, p.wl, p.image_size, p.image_size,
                   tf.rt(), tf.rt(), nullptr, nullptr, a->dev_inter_out, a->dev_union_out, st);
  if (ctx->group_fallback)
    return fail(MLIIS_ERR_STATE, "task-batched execution needs the tensor-core modes (a layer fell back to the "
                                 "single-slot fp32 GEMM kernels)");
  return MLIIS_OK;
}

// a task-batched call on slots slot .. slot + n_group - 1: tensor-core mode, all bound, one layout at a uniform stride
static int group_validate(mliis_ctx* ctx, int32_t slot, int32_t n_group, int64_t stride_bytes) {
  if (n_group <= 1) return MLIIS_OK;
  if (ctx->cfg.gemm_mode == MLIIS_GEMM_FP32) return fail(MLIIS_ERR_ARG, "task-batched execution needs a tensor-core gemm_mode");
  if (slot + n_group > (int)ctx->slots.size()) return fail(MLIIS_ERR_ARG, "group exceeds n_slots");
  if (stride_bytes <= 0 || (stride_bytes & 255)) return fail(MLIIS_ERR_ARG, "group stride must be a positive multiple of 256");
  const Slot& s0 = ctx->slots[slot];
  for (int k = 1; k < n_group; ++k) {
    const Slot& sk = ctx->slots[slot + k];
    if (!sk.state || !sk.ws) return fail(MLIIS_ERR_STATE, "slot %d not bound", slot + k);
    if ((const char*)sk.state - (const char*)s0.state != (ptrdiff_t)k * stride_bytes ||
        (const char*)sk.ws - (const char*)s0.ws != (ptrdiff_t)k * stride_bytes)
      return fail(MLIIS_ERR_ARG, "slots %d..%d are not laid out at the uniform group stride", slot, slot + n_group - 1);
  }
  return MLIIS_OK;
}

static int task_validate(mliis_ctx* ctx, int32_t slot, const mliis_task_args* a) {
  if (!a) return fail(MLIIS_ERR_ARG, "null args");
  int rc = validate(ctx, slot, a->batch);
  if (rc) return rc;
  if (a->n_query < 1 || a->n_query > ctx->plan.maxB) return fail(MLIIS_ERR_ARG, "n_query out of range");
  if (a->n_steps < 0) return fail(MLIIS_ERR_ARG, "n_steps < 0");
  if (!a->dev_init_state || !a->dev_images || !a->dev_labels || !a->dev_batch_index || !a->dev_lr || !a->dev_query_index)
    return fail(MLIIS_ERR_ARG, "null task argument");
  return group_validate(ctx, slot, a->n_group, a->group_stride_bytes);
}

int mliis_adapt_eval_task(mliis_ctx* ctx, int32_t slot, const mliis_task_args* a, void* stream) {
  int rc = task_validate(ctx, slot, a);
  if (rc) return rc;
  rc = task_body(ctx, slot, a, (cudaStream_t)stream);
  if (rc) return rc;
  return check_cuda("adapt_eval_task");
}

int mliis_task_graph_capture(mliis_ctx* ctx, int32_t slot, const mliis_task_args* a, void* stream) {
  int rc = task_validate(ctx, slot, a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (st == nullptr) return fail(MLIIS_ERR_ARG, "graph capture needs a non-default stream");
  Slot& sl = ctx->slots[slot];
  if (sl.graph_exec) { cudaGraphExecDestroy(sl.graph_exec); sl.graph_exec = nullptr; }
  if (sl.graph) { cudaGraphDestroy(sl.graph); sl.graph = nullptr; }
  for (int k = 0; k < (a->n_group > 1 ? a->n_group : 1); ++k) {   // memsets outside the graph: replays do not repeat them
    Slot& sk = ctx->slots[slot + k];
    if (!sk.grads_zeroed) {
      cudaMemsetAsync(sk.ws + ctx->plan.grads, 0, ctx->plan.n_theta * sizeof(float), st);
      sk.grads_zeroed = true;
    }
  }
  cudaStreamSynchronize(st);
  const unsigned long long before = g_kernel_launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    return fail(MLIIS_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(cudaGetLastError()));
  rc = task_body(ctx, slot, a, st);
  cudaError_t e = cudaStreamEndCapture(st, &sl.graph);
  sl.graph_launches = g_kernel_launches - before;
  g_kernel_launches = before;   // captured, not launched
  if (rc) return rc;
  if (e != cudaSuccess || !sl.graph) return fail(MLIIS_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&sl.graph_exec, sl.graph, 0);
  if (e != cudaSuccess) return fail(MLIIS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  return MLIIS_OK;
}

int mliis_task_graph_launch(mliis_ctx* ctx, int32_t slot, void* stream) {
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) return fail(MLIIS_ERR_ARG, "bad ctx/slot");
  Slot& sl = ctx->slots[slot];
  if (!sl.graph_exec) return fail(MLIIS_ERR_STATE, "no captured graph for slot %d", slot);
  cudaError_t e = cudaGraphLaunch(sl.graph_exec, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(MLIIS_ERR_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(e));
  g_kernel_launches += sl.graph_launches;
  return MLIIS_OK;
}

uint64_t mliis_launch_count(void) { return g_kernel_launches; }

// CRC-32C, slicing-by-8 (reflected polynomial 0x82F63B78); host code
uint32_t mliis_crc32c(const void* data, uint64_t n_bytes, uint32_t crc) {
  static uint32_t T[8][256];
  static std::once_flag once;
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      T[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFFu];
  });
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  while (n_bytes >= 8) {
    uint64_t w;
    std::memcpy(&w, p, 8);                       // little-endian host (x86-64 / aarch64)
    w ^= c;
    c = T[7][w & 0xFF] ^ T[6][(w >> 8) & 0xFF] ^ T[5][(w >> 16) & 0xFF] ^ T[4][(w >> 24) & 0xFF] ^
        T[3][(w >> 32) & 0xFF] ^ T[2][(w >> 40) & 0xFF] ^ T[1][(w >> 48) & 0xFF] ^ T[0][(w >> 56) & 0xFF];
    p += 8;
    n_bytes -= 8;
  }
  while (n_bytes--) c = T[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
  return ~c;
}

int mliis_delta_accumulate(mliis_ctx* ctx, float* dsum, const float* ta, const float* tb, int32_t first, void* stream) {
  if (!ctx || !dsum || !ta || !tb) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx");
  delta_accumulate(dsum, ta, tb, ctx->plan.n_theta, first, (cudaStream_t)stream);
  return check_cuda("delta_accumulate");
}
int mliis_meta_apply(mliis_ctx* ctx, float* theta, const float* dsum, float scale, void* stream) {
  if (!ctx || !theta || !dsum) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx");
  meta_apply(theta, dsum, scale, ctx->plan.n_theta, (cudaStream_t)stream);
  return check_cuda("meta_apply");
}

// ---- meta-update exchange: [sum of task deltas | BN moving statistics | #contributors] ----
namespace {
struct RowPtrs { const float* p[32]; };
// out[i] = sum_r rows[r][i] (i < P) ; out[P + j] = sum_r bn_r[j] (j < 2 n_bn) ; out[P + 2 n_bn] = n_rows
__global__ void meta_reduce_kernel(float* __restrict__ out, RowPtrs dsum, RowPtrs bn, int n_rows, int64_t P, int64_t nbn2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    float s = 0.f;
    for (int r = 0; r < n_rows; ++r) s += dsum.p[r][i];
    out[i] = s;
  } else if (i < P + nbn2) {
    float s = 0.f;
    for (int r = 0; r < n_rows; ++r) s += bn.p[r][i - P];
    out[i] = s;
  } else if (i == P + nbn2) {
    out[i] = (float)n_rows;
  }
}
// theta += scale * buf[:P] ; every listed slot's BN statistics <- buf[P:P+2 n_bn] / buf[P + 2 n_bn]
struct WPtrs { float* p[32]; };
__global__ void meta_finish_kernel(float* __restrict__ theta, const float* __restrict__ buf, float scale, WPtrs bn,
                                   int n_slots, int64_t P, int64_t nbn2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    theta[i] = fmaf(scale, buf[i], theta[i]);
  } else if (i < P + nbn2) {
    const float cnt = buf[P + nbn2];
    if (cnt > 0.f) {
      const float v = buf[i] / cnt;
      for (int s = 0; s < n_slots; ++s) bn.p[s][i - P] = v;
    }
  }
}

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch.distributed already loaded, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (h) {
      api.lib = h;
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) api.lib = nullptr;
    }
  }
  return api;
}
}  // namespace

int64_t mliis_meta_buffer_floats(const mliis_ctx* c) {
  return c ? c->plan.n_theta + 2 * (int64_t)c->plan.n_bn_ch + 4 : -1;
}

int mliis_comm_unique_id(uint8_t* out128) {
  if (!out128) return fail(MLIIS_ERR_ARG, "null argument");
  NcclApi& n = nccl();
  if (!n.lib) return fail(MLIIS_ERR_STATE, "libnccl.so.2 not found");
  NcclId id;
  int rc = n.GetUniqueId(&id);
  if (rc) return fail(MLIIS_ERR_CUDA, "ncclGetUniqueId: %s", n.GetErrorString ? n.GetErrorString(rc) : "?");
  memcpy(out128, id.internal, 128);
  return MLIIS_OK;
}

int mliis_comm_init(mliis_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world) {
  if (!ctx || !id128) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx (no sm_100 device): there is no CPU fallback");
  if (world < 1 || rank < 0 || rank >= world) return fail(MLIIS_ERR_ARG, "bad rank / world");
  if (ctx->nccl_comm) return fail(MLIIS_ERR_STATE, "communicator already initialised");
  if (world == 1) { ctx->comm_rank = 0; ctx->comm_world = 1; return MLIIS_OK; }
  NcclApi& n = nccl();
  if (!n.lib) return fail(MLIIS_ERR_STATE, "libnccl.so.2 not found");
  NcclId id;
  memcpy(id.internal, id128, 128);
  cudaSetDevice(ctx->device);
  int rc = n.CommInitRank(&ctx->nccl_comm, world, id, rank);
  if (rc) { ctx->nccl_comm = nullptr; return fail(MLIIS_ERR_CUDA, "ncclCommInitRank: %s", n.GetErrorString ? n.GetErrorString(rc) : "?"); }
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return MLIIS_OK;
}

int mliis_comm_destroy(mliis_ctx* ctx) {
  if (!ctx) return MLIIS_OK;
  if (ctx->nccl_comm) { nccl().CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  ctx->comm_world = 1;
  ctx->comm_rank = 0;
  return MLIIS_OK;
}

int mliis_allreduce_delta(mliis_ctx* ctx, float* buf, int64_t count, void* stream) {
  if (!ctx || !buf) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx");
  if (count < 0) return fail(MLIIS_ERR_ARG, "count < 0");
  if (!ctx->nccl_comm) return MLIIS_OK;            // single process: the local sum is the global sum
  int rc = nccl().AllReduce(buf, buf, (size_t)count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, ctx->nccl_comm, (cudaStream_t)stream);
  if (rc) return fail(MLIIS_ERR_CUDA, "ncclAllReduce: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  return MLIIS_OK;
}

int mliis_meta_reduce(mliis_ctx* ctx, float* buf, const float* dsum_rows, int64_t row_stride, int32_t first_slot,
                      int32_t n_rows, void* stream) {
  if (!ctx || !buf) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx");
  if (n_rows < 0 || n_rows > 32) return fail(MLIIS_ERR_ARG, "n_rows must be in [0, 32]");
  if (n_rows > 0 && !dsum_rows) return fail(MLIIS_ERR_ARG, "null rows");
  if (first_slot < 0 || first_slot + n_rows > (int)ctx->slots.size()) return fail(MLIIS_ERR_ARG, "slot range");
  const Plan& p = ctx->plan;
  RowPtrs d{}, b{};
  for (int r = 0; r < n_rows; ++r) {
    if (!ctx->slots[first_slot + r].state) return fail(MLIIS_ERR_STATE, "slot %d not bound", first_slot + r);
    d.p[r] = dsum_rows + (size_t)r * row_stride;
    b.p[r] = ctx->slots[first_slot + r].state + p.n_theta;
  }
  const int64_t n = p.n_theta + 2 * (int64_t)p.n_bn_ch + 1;
  MLIIS_COUNT(), meta_reduce_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, (cudaStream_t)stream>>>(buf, d, b, n_rows, p.n_theta,
                                                                                               2 * (int64_t)p.n_bn_ch);
  return check_cuda("meta_reduce");
}

int mliis_meta_finish(mliis_ctx* ctx, float* theta, const float* buf, float scale, int32_t first_slot, int32_t n_slots,
                      void* stream) {
  if (!ctx || !theta || !buf) return fail(MLIIS_ERR_ARG, "null argument");
  if (ctx->device < 0) return fail(MLIIS_ERR_DEVICE, "table-only ctx");
  if (n_slots < 0 || n_slots > 32 || first_slot < 0 || first_slot + n_slots > (int)ctx->slots.size())
    return fail(MLIIS_ERR_ARG, "slot range");
  const Plan& p = ctx->plan;
  WPtrs b{};
  for (int r = 0; r < n_slots; ++r) {
    if (!ctx->slots[first_slot + r].state) return fail(MLIIS_ERR_STATE, "slot %d not bound", first_slot + r);
    b.p[r] = ctx->slots[first_slot + r].state + p.n_theta;
  }
  const int64_t n = p.n_theta + 2 * (int64_t)p.n_bn_ch;
  MLIIS_COUNT(), meta_finish_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, (cudaStream_t)stream>>>(theta, buf, scale, b, n_slots,
                                                                                               p.n_theta, 2 * (int64_t)p.n_bn_ch);
  return check_cuda("meta_finish");
}

// ---- per-kernel entry points ----
static int require_sm100() {
  // cudaGetDeviceProperties costs milliseconds: query the one attribute and cache it per device
  static int cached_major[64];
  static bool cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(MLIIS_ERR_DEVICE, "no CUDA device (no CPU fallback exists)");
  if (dev < 0 || dev >= 64) return fail(MLIIS_ERR_DEVICE, "bad device ordinal");
  if (!cached[dev]) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return fail(MLIIS_ERR_DEVICE, "cannot query the device (no CPU fallback exists)");
    cached_major[dev] = major;
    cached[dev] = true;
  }
  if (cached_major[dev] != 10) return fail(MLIIS_ERR_DEVICE, "not an sm_100 device (no fallback)");
  return MLIIS_OK;
}

// Task-batched per-kernel calls: mliis_kernel_group(n, stride) makes every following per-kernel entry point of this
// thread launch once for n slot copies laid out `stride` bytes apart (every pointer argument is slot 0's).
int mliis_kernel_group(int32_t n_group, int64_t group_stride_bytes) {
  if (n_group < 1 || n_group > 1024) return fail(MLIIS_ERR_ARG, "n_group must be in [1, 1024]");
  if (n_group > 1 && (group_stride_bytes <= 0 || (group_stride_bytes & 15))) return fail(MLIIS_ERR_ARG, "bad group stride");
  t_kernel_group = ZGroup{n_group, n_group > 1 ? (long long)(group_stride_bytes / 4) : 0};
  return MLIIS_OK;
}
#define KERNEL_GROUP() ZScope zscope_(t_kernel_group.nz, t_kernel_group.zs)

int mliis_dwconv_fwd(const float* x, const float* w, float* y, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k,
                     int32_t stride, const float* bn_a, const float* bn_b, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  KERNEL_GROUP();
  if ((k != 3 && k != 5) || (stride != 1 && stride != 2) || C % 4) return fail(MLIIS_ERR_ARG, "unsupported depthwise shape");
  int pt, pb;
  same_pad(H, k, stride, 1, &pt, &pb);
  int pl, pr;
  same_pad(W, k, stride, 1, &pl, &pr);
  dw_fwd(x, bn_a, bn_b, w, y, B, H, W, C, k, stride, (H + stride - 1) / stride, (W + stride - 1) / stride, pt, pl,
         (cudaStream_t)stream);
  return check_cuda("dwconv_fwd");
}

int mliis_gemm_nn(const float* a, const float* w, float* c, int32_t M, int32_t K, int32_t N, int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (K % 8 || N % 4) return fail(MLIIS_ERR_ARG, "K must be a multiple of 8 and N of 4");
  if (mode != MLIIS_GEMM_FP32) {
    if (!tc_supported(0, 0, K, N)) return fail(MLIIS_ERR_ARG, "shape not supported by the tcgen05 path");
    float* wt = nullptr;
    const int split = mode == MLIIS_GEMM_TF32X3 ? 3 : 1;
    if (cudaMalloc(&wt, (size_t)2 * K * N * sizeof(float)) != cudaSuccess) return fail(MLIIS_ERR_CUDA, "alloc");
    tc_prep_weights(w, wt, 1, K, N, 0, split, (cudaStream_t)stream);
    bool ok = tc_conv(a, K, wt, nullptr, c, N, 0, M, 1, 1, 1, K, 1, 1, N, 0, split, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(wt);
    if (!ok) return fail(MLIIS_ERR_CUDA, "tc_conv setup failed (tensor map encode)");
    return check_cuda("gemm_nn(tc)");
  }
  gemm_nn(plainA(a, K), w, nullptr, c, N, M, K, N, M, 0, (cudaStream_t)stream);
  return check_cuda("gemm_nn");
}

int mliis_conv3x3_fwd(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t H, int32_t W,
                      int32_t Cin, int32_t Cout, int32_t dilation, int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (Cin % 8 || Cout % 4) return fail(MLIIS_ERR_ARG, "Cin must be a multiple of 8 and Cout of 4");
  if (mode != MLIIS_GEMM_FP32) {
    if (!tc_supported(1, W, Cin, Cout)) return fail(MLIIS_ERR_ARG, "shape not supported by the tcgen05 path");
    float* wt = nullptr;
    const int split = mode == MLIIS_GEMM_TF32X3 ? 3 : 1;
    if (cudaMalloc(&wt, (size_t)2 * 9 * Cin * Cout * sizeof(float)) != cudaSuccess) return fail(MLIIS_ERR_CUDA, "alloc");
    tc_prep_weights(w, wt, 9, Cin, Cout, 0, split, (cudaStream_t)stream);
    bool ok = tc_conv(x, Cin, wt, bias, y, Cout, 1, B * H * W, B, H, W, Cin, 9, dilation, Cout, 0, split,
                      (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(wt);
    if (!ok) return fail(MLIIS_ERR_CUDA, "tc_conv setup failed (tensor map encode)");
    return check_cuda("conv3x3_fwd(tc)");
  }
  gemm_nn(convA(x, Cin, H, W, Cin, dilation), w, bias, y, Cout, B * H * W, 9 * Cin, Cout, H * W, 0, (cudaStream_t)stream);
  return check_cuda("conv3x3_fwd");
}

int mliis_tc_prep_weights(const float* w, float* wt, int32_t taps, int32_t Cin, int32_t Cout, int32_t dgrad,
                          int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  KERNEL_GROUP();
  if (mode == MLIIS_GEMM_FP32) return fail(MLIIS_ERR_ARG, "mode must be a tensor-core mode");
  tc_prep_weights(w, wt, taps, Cin, Cout, dgrad, mode == MLIIS_GEMM_TF32X3 ? 3 : 1, (cudaStream_t)stream);
  return check_cuda("tc_prep_weights");
}

int mliis_tc_conv(const float* x, const float* wt, const float* bias, float* y, int32_t B, int32_t H, int32_t W,
                  int32_t Cin, int32_t Cout, int32_t taps, int32_t dilation, int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  KERNEL_GROUP();
  if (mode == MLIIS_GEMM_FP32 || (taps != 1 && taps != 9)) return fail(MLIIS_ERR_ARG, "bad mode / taps");
  const int conv = taps == 9;
  if (!tc_supported(conv, W, Cin, Cout)) return fail(MLIIS_ERR_ARG, "shape not supported by the tcgen05 path");
  if (!tc_conv(x, Cin, wt, bias, y, Cout, conv, B * H * W, B, H, W, Cin, taps, dilation, Cout, 0,
               mode == MLIIS_GEMM_TF32X3 ? 3 : 1, (cudaStream_t)stream))
    return fail(MLIIS_ERR_CUDA, "tc_conv setup failed (tensor map encode)");
  return check_cuda("tc_conv");
}

// MBConv project conv (efficientnet_model.py:271-273 after :225-232, :266): y[M, Cout] = (swish(bn_a*x + bn_b) * gate[img]) * W.
// The normalised, activated, SE-gated tensor never exists in HBM: the prologue runs in the operand loader.
int mliis_tc_project_conv(const float* x, const float* wt, const float* bn_a, const float* bn_b, const float* gate, float* y,
                          int32_t B, int32_t HW, int32_t Cin, int32_t Cout, int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!x || !wt || !bn_a || !bn_b || !y) return fail(MLIIS_ERR_ARG, "null argument");
  if (mode == MLIIS_GEMM_FP32) return fail(MLIIS_ERR_ARG, "mode must be a tensor-core mode");
  if (!tc_supported(0, 0, Cin, Cout)) return fail(MLIIS_ERR_ARG, "shape not supported by the tcgen05 path");
  KERNEL_GROUP();
  if (!tc_conv(x, Cin, wt, nullptr, y, Cout, 0, B * HW, B, 1, 1, Cin, 1, 1, Cout, 0, mode == MLIIS_GEMM_TF32X3 ? 3 : 1,
               (cudaStream_t)stream, bn_a, bn_b, gate, HW))
    return fail(MLIIS_ERR_CUDA, "tc_conv setup failed (tensor map encode)");
  return check_cuda("tc_project_conv");
}

int mliis_tc_wgrad(const float* a, const float* g, float* dw, int32_t B, int32_t H, int32_t W, int32_t Cin,
                   int32_t Cout, int32_t taps, int32_t dilation, int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (mode == MLIIS_GEMM_FP32 || (taps != 1 && taps != 9)) return fail(MLIIS_ERR_ARG, "bad mode / taps");
  const int conv = taps == 9, M = B * H * W;
  if (!tc_wgrad_supported(conv, W, Cin, Cout)) return fail(MLIIS_ERR_ARG, "shape not supported by the tcgen05 wgrad");
  float* scratch = nullptr;
  if (cudaMalloc(&scratch, tc_wgrad_scratch(conv, M, B, H, W, Cin, Cout, taps) * sizeof(float)) != cudaSuccess)
    return fail(MLIIS_ERR_CUDA, "alloc");
  bool ok = tc_wgrad(a, Cin, g, Cout, dw, scratch, conv, M, B, H, W, Cin, taps, dilation, Cout,
                     mode == MLIIS_GEMM_TF32X3 ? 3 : 1, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(scratch);
  if (!ok) return fail(MLIIS_ERR_CUDA, "tc_wgrad setup failed (tensor map encode)");
  return check_cuda("tc_wgrad");
}

int mliis_bilinear_fwd(const float* x, float* y, int32_t B, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout,
                       int32_t C, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (C % 4) return fail(MLIIS_ERR_ARG, "C must be a multiple of 4");
  Tab ty, tx;
  if (!make_tab(nullptr, Hin, Hout, &ty) || !make_tab(nullptr, Win, Wout, &tx)) return fail(MLIIS_ERR_CUDA, "table alloc");
  bilinear_fwd(x, C, y, C, B, Hin, Win, Hout, Wout, C, ty.rt(), tx.rt(), (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  for (Tab* t : {&ty, &tx}) { cudaFree(t->lo); cudaFree(t->hi); cudaFree(t->lerp); cudaFree(t->g_lo); cudaFree(t->g_hi); }
  return check_cuda("bilinear_fwd");
}

int mliis_adam_step(float* theta, float* v, const float* grad, int64_t n, int64_t n_l2, float lr, float beta2_power,
                    float l2_coef, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!theta || !v || !grad || n < 1) return fail(MLIIS_ERR_ARG, "bad argument");
  if (((uintptr_t)theta | (uintptr_t)v | (uintptr_t)grad) & 15) return fail(MLIIS_ERR_ARG, "buffers must be 16-byte aligned");
  KERNEL_GROUP();
  adam_step(theta, v, grad, n, n_l2, nullptr, nullptr, l2_coef, 0, (cudaStream_t)stream, lr, beta2_power);
  return check_cuda("adam_step");
}

// ---- HBM-bound kernels alone (unit tests / roofline micro-benchmarks) ----
int64_t mliis_kernel_scratch_floats(int32_t B, int32_t H, int32_t W, int32_t C) {
  // enough for any of the entry points below on a [B,H,W,C] tensor
  int64_t n = (int64_t)rc_num_chunks(B * H * W, C) * 2 * C + 2 * 1024;
  n = std::max<int64_t>(n, (int64_t)dw_wgrad_blocks(B, H, W, 1) * 25 * C);
  n = std::max<int64_t>(n, (int64_t)B * rc_num_img_chunks(H * W, C) * C + (int64_t)B * 2 * C + 64);
  n = std::max<int64_t>(n, (int64_t)B * 32 * 4 + 148 + 64);
  return n + 1024;
}

int mliis_dwconv_bwd(const float* x, const float* bn_a, const float* bn_b, const float* w, const float* dy, float* dx,
                     float* dw, float* scratch, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                     void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if ((k != 3 && k != 5) || (stride != 1 && stride != 2) || C % 4) return fail(MLIIS_ERR_ARG, "unsupported depthwise shape");
  if (!x || !w || !dy || !dx || !dw || !scratch) return fail(MLIIS_ERR_ARG, "null argument");
  KERNEL_GROUP();
  int pt, pb, pl, pr;
  same_pad(H, k, stride, 1, &pt, &pb);
  same_pad(W, k, stride, 1, &pl, &pr);
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  dw_bwd_weight(x, bn_a, bn_b, dy, scratch, dw, B, H, W, C, k, stride, Ho, Wo, pt, pl, (cudaStream_t)stream);
  dw_bwd_data(dy, w, dx, B, H, W, C, k, stride, Ho, Wo, pt, pl, (cudaStream_t)stream);
  return check_cuda("dwconv_bwd");
}

// The clustered reductions elect their finalizing CTA with a ticket word that must start at zero and resets itself.  The
// engine keeps it in the workspace (zeroed at mliis_slot_bind); a per-kernel entry point takes it from the caller's scratch
// and clears it the first time it sees that scratch pointer (every slot copy of the current kernel group).
static unsigned* entry_ticket(float* where, cudaStream_t st) {
  static thread_local const void* cleared = nullptr;
  static thread_local int cleared_nz = 0;
  if (cleared != where || cleared_nz != MLIIS_NZ) {
    for (int z = 0; z < MLIIS_NZ; ++z) cudaMemsetAsync(where + (size_t)z * MLIIS_ZS, 0, 16 * sizeof(float), st);
    cleared = where;
    cleared_nz = MLIIS_NZ;
  }
  return reinterpret_cast<unsigned*>(where);
}

// train-mode BN forward bookkeeping of one layer: batch statistics of x [M,C] -> stats = [mean | rstd | a | b] (4*C),
// EMA of the moving statistics (utils.py:111-134).  The normalise + swish itself is fused into the consumers.
int mliis_bn_stats_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                       float* stats, float* scratch, int32_t M, int32_t C, int32_t fused, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (C % 4 || C > 1024 || M < 2) return fail(MLIIS_ERR_ARG, "C must be a multiple of 4 and <= 1024");
  if (!x || !gamma || !beta || !moving_mean || !moving_var || !stats || !scratch) return fail(MLIIS_ERR_ARG, "null argument");
  KERNEL_GROUP();
  unsigned* ticket = entry_ticket(scratch + (size_t)rc_num_chunks(M, C) * 2 * C + 2 * C, (cudaStream_t)stream);
  bn_stats_finalize(x, C, M, C, false, scratch, ticket, gamma, beta, moving_mean, moving_var, 1, fused, stats, stats + C,
                    stats + 2 * C, stats + 3 * C, (cudaStream_t)stream);
  return check_cuda("bn_stats_fwd");
}

// backward of y = swish(BN(x)) (the MBConv expand / stem sites): dgamma, dbeta, dx   (reduce + finalize + apply)
int mliis_bn_swish_bwd(const float* x, const float* g, float* dx, const float* stats, const float* gamma, float* dgamma,
                       float* dbeta, float* scratch, int32_t M, int32_t C, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (C % 4 || C > 1024) return fail(MLIIS_ERR_ARG, "C must be a multiple of 4 and <= 1024");
  if (!x || !g || !dx || !stats || !gamma || !dgamma || !dbeta || !scratch) return fail(MLIIS_ERR_ARG, "null argument");
  KERNEL_GROUP();
  BnBwdArgs a{};
  a.x = x; a.ldx = C; a.g = g; a.ldg = C; a.dx = dx; a.lddx = C; a.M = M; a.C = C; a.HW = M;
  a.mean = stats; a.rstd = stats + C; a.a = stats + 2 * C; a.b = stats + 3 * C; a.gamma = gamma;
  a.partials = scratch; a.k = scratch + (size_t)rc_num_chunks(M, C) * 2 * C; a.dgamma = dgamma; a.dbeta = dbeta;
  a.ticket = entry_ticket(a.k + 2 * C, (cudaStream_t)stream);
  bn_bwd(BN_SWISH, a, (cudaStream_t)stream);
  return check_cuda("bn_swish_bwd");
}

// squeeze-excite forward (efficientnet_model.py:238-251) on the pre-BN depthwise output x [B,HW,C]: pooled
// swish(a*x+b) -> FC -> swish -> FC -> sigmoid = gate [B,C] (the multiply lives in the project conv's loader)
int mliis_se_fwd(const float* x, const float* bn_a, const float* bn_b, const float* w1, const float* b1, const float* w2,
                 const float* b2, float* pool, float* hidpre, float* gate, float* scratch, int32_t B, int32_t HW,
                 int32_t C, int32_t Cr, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (C % 4 || C > 1024 || Cr < 1) return fail(MLIIS_ERR_ARG, "bad SE shape");
  if (!x || !bn_a || !bn_b || !w1 || !b1 || !w2 || !b2 || !pool || !hidpre || !gate || !scratch) return fail(MLIIS_ERR_ARG, "null argument");
  KERNEL_GROUP();
  se_pool(x, C, bn_a, bn_b, B, HW, C, scratch, (cudaStream_t)stream);
  se_fc_fwd(scratch, rc_num_img_chunks(HW, C), B, HW, C, Cr, w1, b1, w2, b2, pool, hidpre, gate, (cudaStream_t)stream);
  return check_cuda("se_fwd");
}

// fused loss of the binary head (efficientlab.py:294-327): bilinear upsample of the low-res logits, softmax-CE
// (+ label smoothing) - ln(dice), per-image soft-IoU sums, and the gradient w.r.t. the FULL-resolution logits.
int mliis_softmax_ce_iou(const float* z_lo, const float* labels, float* p1, float* dz_hi, float* scratch, float* loss_out,
                         int32_t B, int32_t h, int32_t w, int32_t H, int32_t W, int32_t dice, float label_smoothing,
                         void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!z_lo || !labels || !p1 || !dz_hi || !scratch) return fail(MLIIS_ERR_ARG, "null argument");
  static std::vector<std::pair<std::pair<int, int>, Tab>> cache;     // resize tables per (in, out), kept for the process
  auto tab = [&](int n_in, int n_out) -> const Tab* {
    for (auto& e : cache) if (e.first == std::make_pair(n_in, n_out)) return &e.second;
    Tab t;
    if (!make_tab(nullptr, n_in, n_out, &t)) return nullptr;
    cache.push_back({{n_in, n_out}, t});
    return &cache.back().second;
  };
  const Tab* ty = tab(h, H);
  const Tab* tx = tab(w, W);
  if (!ty || !tx) return fail(MLIIS_ERR_CUDA, "table alloc");
  KERNEL_GROUP();
  LossArgs la{};
  la.z_lo = z_lo; la.labels = labels; la.index = nullptr; la.B = B; la.h = h; la.w = w; la.H = H; la.W = W;
  la.ty = ty->rt(); la.tx = tx->rt(); la.dice = dice; la.label_smoothing = label_smoothing;
  la.p1 = p1; la.partials = scratch; la.coef = scratch + (size_t)B * 32 * 4 + 148 + 64; la.dz_hi = dz_hi; la.loss_out = loss_out;
  la.theta = nullptr; la.n_l2 = 0; la.l2_coef = 0.f;
  loss_fwd_bwd(la, (cudaStream_t)stream);
  return check_cuda("softmax_ce_iou");
}

// conv2d_2 of an RSD module with the image-pooling branch FOLDED (k_pool.cu): y = conv3x3(x[:, :Cin]) + bias +
// bias9[img][border class], bias9 from pooled [B,Cp] and the kernel rows Cin..Cin+Cp of every tap.  dev_wt: the
// operand prepared by mliis_tc_prep_weights_sub (first Cin of the Cin+Cp input channels).  Two launches.
int mliis_tc_prep_weights_sub(const float* w, float* wt, int32_t taps, int32_t Cin, int32_t Cs, int32_t Cout, int32_t dgrad,
                              int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (mode == MLIIS_GEMM_FP32 || Cs < Cin) return fail(MLIIS_ERR_ARG, "bad mode / Cs");
  KERNEL_GROUP();
  tc_prep_weights(w, wt, taps, Cin, Cout, dgrad, mode == MLIIS_GEMM_TF32X3 ? 3 : 1, (cudaStream_t)stream, Cs);
  return check_cuda("tc_prep_weights_sub");
}
int mliis_rsd_conv2_fwd(const float* x, int32_t ldx, const float* pooled, const float* w_hwio, const float* wt,
                        const float* bias, float* bias9, float* y, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cp,
                        int32_t Cout, int32_t mode, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (mode == MLIIS_GEMM_FP32) return fail(MLIIS_ERR_ARG, "mode must be a tensor-core mode");
  if (!x || !pooled || !w_hwio || !wt || !bias9 || !y) return fail(MLIIS_ERR_ARG, "null argument");
  if (Cin % 4 || Cout % 4 || ldx % 4 || H < 2 || W < 2) return fail(MLIIS_ERR_ARG, "bad shape");
  KERNEL_GROUP();
  pool_bias9(pooled, Cp, w_hwio, Cin + Cp, Cin, Cp, Cout, B, bias9, (cudaStream_t)stream);
  if (!tc_conv(x, ldx, wt, bias, y, Cout, 1, B * H * W, B, H, W, Cin, 9, 1, Cout, 0, mode == MLIIS_GEMM_TF32X3 ? 3 : 1,
               (cudaStream_t)stream, nullptr, nullptr, nullptr, 0, bias9))
    return fail(MLIIS_ERR_ARG, "shape not supported by the folded tcgen05 path");
  return check_cuda("rsd_conv2_fwd");
}

int mliis_tc_peak_tf32(int32_t iters, double* tflops_out, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!tflops_out || iters < 1) return fail(MLIIS_ERR_ARG, "bad argument");
  *tflops_out = tc_peak_tf32(iters, (cudaStream_t)stream);
  return check_cuda("tc_peak_tf32");
}

int mliis_tc_mma_rate(int32_t iters, int32_t n, int32_t pattern, int32_t shift_rows, double* clk_per_kstep_out, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!clk_per_kstep_out || iters < 1) return fail(MLIIS_ERR_ARG, "bad argument");
  const double v = tc_mma_rate(iters, n, pattern, shift_rows, (cudaStream_t)stream);
  if (v < 0) return fail(MLIIS_ERR_ARG, "bad MMA shape / pattern");
  *clk_per_kstep_out = v;
  return check_cuda("tc_mma_rate");
}

int mliis_debug_buffer(mliis_ctx* ctx, int32_t slot, const char* name, const float** dev_ptr, int64_t* rows_per_image,
                       int32_t* channels, int32_t* ld) {
  if (!ctx || !name || !dev_ptr) return fail(MLIIS_ERR_ARG, "null argument");
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot].ws) return fail(MLIIS_ERR_STATE, "slot not bound");
  const Plan& p = ctx->plan;
  if (std::strcmp(name, "grads") == 0) {
    *dev_ptr = ctx->slots[slot].ws + p.grads;
    if (rows_per_image) *rows_per_image = 1;
    if (channels) *channels = (int32_t)p.n_theta;
    if (ld) *ld = (int32_t)p.n_theta;
    return MLIIS_OK;
  }
  for (const Plan::Named& n : p.named) {
    if (n.name == name) {
      *dev_ptr = ctx->slots[slot].ws + n.buf.off;
      if (rows_per_image) *rows_per_image = n.buf.HW;
      if (channels) *channels = n.buf.C;
      if (ld) *ld = n.buf.ld;
      return MLIIS_OK;
    }
  }
  return fail(MLIIS_ERR_ARG, "unknown buffer '%s'", name);
}

}  // extern "C"
