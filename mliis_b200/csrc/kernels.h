// Host-side launchers of the mliis_b200 kernels.  All pointers are device pointers; all tensors are
// fp32 NHWC viewed as [rows = B*H*W, C] with an explicit row stride `ld` (floats) so that producers can
// write straight into channel slices of concat buffers (the reference's tf.concat is never materialised
// by a copy of its own).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mliis {

// number of kernels launched by this library in this process (bench.py reports it as gpu_launches)
extern unsigned long long g_kernel_launches;
// MLIIS_SKIP=<comma-separated substrings of launcher function names> drops those launches (bottleneck experiments
// only: results are garbage); unset in every normal run.
bool skip_launch(const char* launcher);
// usage: `MLIIS_COUNT(), kernel<<<grid, block, smem, stream>>>(...);`  (a one-trip for statement: safe inside un-braced
// if / else bodies, unlike an if / else macro)
#define MLIIS_COUNT() \
  for (bool mliis_go_ = !::mliis::skip_launch(__func__); mliis_go_; mliis_go_ = false) ++::mliis::g_kernel_launches

// Task-batched launches (common.cuh): the engine sets the group for the calling thread; every launcher multiplies
// grid.z by nz and passes zs (floats between consecutive slots) to its kernel.  Default {1, 0} = one slot.
struct ZGroup { int nz; long long zs; };
ZGroup& zgroup();
struct ZScope {
  ZGroup saved;
  ZScope(int nz, long long zs) : saved(zgroup()) { zgroup() = ZGroup{nz, zs}; }
  ~ZScope() { zgroup() = saved; }
};
#define MLIIS_NZ (::mliis::zgroup().nz)
// How many slots share the launch when a launcher sizes its reduction partials (wgrad pixel-range splits, row chunks of
// the BN / SE sums): nz, so that a task-batched launch keeps about the same number of CTAs - and nz times fewer
// partials per slot - as a single-slot one.  The summation tree then depends on the group size (results stay
// deterministic for a given group size).  MLIIS_GROUP_CANONICAL=1 (read at every call) returns 1: the single-slot
// partition for every group size, which makes task-batched launches BIT-IDENTICAL to single-slot ones (tests).
int partition_nz();
#define MLIIS_ZS (::mliis::zgroup().zs)

// ---------------- row-channel kernels (k_rowchan.cu) : HBM-bound ----------------
enum BnVar { BN_PLAIN = 0, BN_SWISH = 1, BN_SWISH_SE = 2, BN_DEC = 3 };

int rc_num_chunks(int M, int C);              // #row chunks used by the whole-tensor reductions
int rc_num_img_chunks(int HW, int C);         // #row chunks per image used by the per-image reductions

// batch statistics of x (pre_swish: of swish(x)) -> partials [G][2][C]
// train-mode batch statistics AND their finalize (mean, rstd, affine coefficients, EMA of the moving statistics) in one
// launch: clusters of 8 row-chunk CTAs reduce through distributed shared memory, the last cluster leader finalizes
void bn_stats_finalize(const float* x, int ld, int M, int C, bool pre_swish, float* partials, unsigned* ticket,
                       const float* gamma, const float* beta, float* mm, float* mv, int ema, int bessel, float* mean,
                       float* rstd, float* a, float* b, cudaStream_t s);
// eval mode coefficients for ALL layers at once: a = gamma*rsqrt(mv+eps), b = beta - mm*a
void bn_eval_coeffs(const float* theta, const int32_t* gamma_idx, const int32_t* beta_idx, const float* moving_mean,
                    const float* moving_var, int n_ch, float* a, float* b, cudaStream_t s);

// decoder: y = a*swish(x)+b (+res)
void dec_bn_apply(const float* x, int ldx, const float* a, const float* b, const float* res, int ldres, float* y,
                  int ldy, int M, int C, cudaStream_t s);
// backbone block output: y = (a*x+b)*dcs[img] + res   (dcs, res nullable)
void block_out(const float* x, int ldx, const float* a, const float* b, const float* dcs, const float* res,
               int ldres, float* y, int ldy, int M, int C, int HW, cudaStream_t s);

// SE squeeze: partial[b][g][C] = sum_rows swish(a*x+b)
void se_pool(const float* x, int ldx, const float* a, const float* b, int B, int HW, int C, float* partial,
             cudaStream_t s);
// SE FCs: pool, hidden pre-activation, gate (efficientnet_model.py:238-251)
void se_fc_fwd(const float* partial, int G, int B, int HW, int C, int Cr, const float* w1, const float* b1,
               const float* w2, const float* b2, float* pool, float* hidpre, float* gate, cudaStream_t s);
// dgate partial[b][g][C] = sum_rows gup * swish(a*x+b)
void se_bwd_reduce(const float* x, int ldx, const float* gup, int ldg, const float* a, const float* b, int B,
                   int HW, int C, float* partial, cudaStream_t s);
void se_fc_bwd(const float* partial, int G, int B, int HW, int C, int Cr, const float* w1, const float* w2,
               const float* pool, const float* hidpre, const float* gate, float* dw1, float* db1, float* dw2,
               float* db2, float* dpool /* [B][C], already / HW */, cudaStream_t s);

struct BnBwdArgs {
  const float* x; int ldx;      // pre-BN tensor
  const float* g; int ldg;      // upstream gradient
  float* dx; int lddx;          // output (may alias g)
  int M, C, HW;
  const float* mean; const float* rstd; const float* a; const float* b; const float* gamma;
  const float* dcs;             // [B] or null (BN_PLAIN)
  const float* gate;            // [B][C] (BN_SWISH_SE)
  const float* dpool;           // [B][C] (BN_SWISH_SE)
  float* partials;              // [G/8][2][C] (one row per cluster of row chunks)
  unsigned* ticket;             // one zero-initialised word per slot (self-resetting): elects the finalizing CTA
  float* k;                     // [2][C]  mean(g), mean(g*xhat)
  float* dgamma; float* dbeta;  // [C] gradient outputs
};
void bn_bwd(int var, const BnBwdArgs& p, cudaStream_t s);  // reduce + finalize + apply

// per-image column sums: out[b][c] = scale * sum_rows x ; and helpers for concat/pool plumbing
void img_colsum(const float* x, int ldx, int B, int HW, int C, float scale, float* partial, float* out, int ldo,
                cudaStream_t s);
void bcast_rows(const float* pimg, int ldp, float* y, int ldy, int B, int HW, int C, cudaStream_t s);
// dst = a + b + pimg[img]  (b, pimg nullable)
void add3(float* dst, int ldd, const float* a, int lda, const float* b, int ldb, const float* pimg, int ldp, int M,
          int C, int HW, cudaStream_t s);

// ---------------- convolution kernels (k_conv.cu) ----------------
// stem: normalise + conv3x3 s2 (efficientlab.py:113-114, efficientnet_model.py:359-366)
void stem_fwd(const float* images, const int32_t* index, const float* w, float* y, int B, int H, int W, int Ho,
              int Wo, int pad_t, int pad_l, cudaStream_t s);
void stem_wgrad(const float* images, const int32_t* index, const float* dy, float* partials, float* dw, int B, int H,
                int W, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s);
int stem_wgrad_blocks(int B, int Ho, int Wo);

// depthwise k x k, stride 1|2, TF SAME.  Input is swish(a*x+b) when a != null (prologue fusion).
void dw_fwd(const float* x, const float* a, const float* b, const float* w, float* y, int B, int H, int W, int C,
            int k, int stride, int Ho, int Wo, int pad_t, int pad_l, cudaStream_t s);
void dw_bwd_data(const float* dy, const float* w, float* dx, int B, int H, int W, int C, int k, int stride, int Ho,
                 int Wo, int pad_t, int pad_l, cudaStream_t s);
int dw_wgrad_blocks(int B, int Ho, int Wo, int stride);
void dw_bwd_weight(const float* x, const float* a, const float* b, const float* dy, float* partials, float* dw,
                   int B, int H, int W, int C, int k, int stride, int Ho, int Wo, int pad_t, int pad_l,
                   cudaStream_t s);

// ---------------- dense contractions, fp32 FFMA (k_gemm.cu) ----------------
struct GemmA {              // A operand description: rows m = (b,y,x) of an NHWC tensor
  const float* ptr; int ld;
  // prologue: v = swish(pa[k]*v + pb[k]) * gate[img*K + k]   (pa null = none; gate null = none)
  const float* pa; const float* pb; const float* gate;
  // implicit 3x3 conv: K = 9*C, tap (ty,tx) reads pixel (y+(ty-1)*dil, x+(tx-1)*dil), zero outside
  int conv;                 // 0: plain [M,K];  1: 3x3 taps
  int H, W, C, dil;
};
// C[M,N] (ldc) = A[M,K] * Wt[K,N] (+bias[n]) (+C if accumulate)
void gemm_nn(const GemmA& A, const float* Wt, const float* bias, float* Cout, int ldc, int M, int K, int N,
             int HW, int accumulate, cudaStream_t s);
// dW[K,N] = sum_m A[m,k] * G[m,n]; conv=1: K = 9*C taps; partial scratch sized by gemm_tn_scratch()
size_t gemm_tn_scratch(int M, int K, int N, int conv);
void gemm_tn(const GemmA& A, const float* G, int ldg, float* dW, float* dbias /* nullable: sum_m G */, float* scratch,
             int M, int K, int N, int HW, cudaStream_t s);
// Wt[N,K] <- W[K,N]   and   3x3: Wf[tap][n][c] <- W[8-tap][c][n]
void transpose_w(const float* w, float* wt, int K, int N, cudaStream_t s);
void flip_transpose_w3x3(const float* w, float* wt, int C, int N, cudaStream_t s);

// ---------------- dense contractions on tcgen05 tensor cores (k_tc.cu) ----------------
// C[m,n] = sum_{tap,c} A[pixel(m)+tap, c] * Wt[n][tap][c] (+bias) (+C); TF32 operands, fp32 TMEM accumulator.
// conv=1: A is NHWC [B,H,W,C] (row stride lda), taps = 9 (3x3, dilation dil); conv=0: A is [M,C], taps = 1.
// Returns false when the shape is not supported (caller falls back to the fp32 FFMA kernels of k_gemm.cu).
bool tc_supported(int conv, int W, int C, int N);
// split = 1: TF32 with round-to-nearest operands; split = 3: 3xTF32 (hi/lo planes, fp32-class accuracy).
// Optional A prologue (conv = 0 only): a <- swish(pa[k]*a + pb[k]) * gate[row / HW][k], applied in shared memory by
// the transform warps (MBConv project conv: BN1 + swish + squeeze-excite gate never touch HBM).
// bias9 (3x3 only): per-image, per-TAP vectors [B][9][N] of the folded pooled branch (k_pool.cu); the kernel sums the taps
// valid in each border class and adds that class bias in the epilogue;
// returns false if the shape does not fit the kernel that implements it (tc_conv3_kernel).
bool tc_conv(const float* A, int lda, const float* Wt, const float* bias, float* out, int ldc, int conv, int M, int B,
             int H, int W, int C, int taps, int dil, int N, int accumulate, int split, cudaStream_t s,
             const float* pa = nullptr, const float* pb = nullptr, const float* gate = nullptr, int HW = 0,
             const float* bias9 = nullptr);
// W[tap][ci][co] (HWIO) -> Wt[co][tap][ci] (dgrad=0)  or  Wt[ci][taps-1-tap][co] (dgrad=1), rn(tf32);
// split == 3 appends the residual plane (wt must hold 2 * taps*Ci*Co floats).  Cs >= Ci: input channels per tap of
// the SOURCE tensor (the operand uses its first Ci: conv2d_2 without the folded pooled channels); 0 = Ci.
void tc_prep_weights(const float* w, float* wt, int taps, int Ci, int Co, int dgrad, int split, cudaStream_t s,
                     int Cs = 0);
// one launch per step: operand (re-layout + TF32 hi/lo split) of every dense layer, forward and dgrad
struct TcPrepJob { long long w_off, dst; int taps, Ci, Co, dgrad, split, Cs; };
void tc_prep_all(const float* theta, float* wcache, const TcPrepJob* dev_jobs, int n_jobs, cudaStream_t s);

// measured kind::tf32 cta_group::1 tensor-pipe peak (TFLOP/s): operands resident in shared memory, no loads
double tc_peak_tf32(int iters, cudaStream_t s);
double tc_mma_rate(int iters, int n, int pattern, int shift_rows, cudaStream_t s);

// wgrad on tensor cores: dW[taps*C, N] = sum_pixels A[pixel+tap, c] * G[pixel, n]  (A, G channel-contiguous)
bool tc_wgrad_supported(int conv, int W, int C, int N);
size_t tc_wgrad_scratch(int conv, int M, int B, int H, int W, int C, int N, int taps);
// dw_tap_stride: floats between consecutive taps of dW (0 = dense taps, C*N)
bool tc_wgrad(const float* A, int lda, const float* G, int ldg, float* dW, float* scratch, int conv, int M, int B, int H,
              int W, int C, int taps, int dil, int N, int split, cudaStream_t s, const float* pa = nullptr,
              const float* pb = nullptr, const float* gate = nullptr, int HW = 0, int dw_tap_stride = 0);

// ---------------- folded image-pooling branch of the RSD decoder (k_pool.cu) ----------------
// w_hwio: conv2d_2 kernel [9][Cs][D]; the pooled channels are rows c_first .. c_first+Cp of every tap.
void pool_bias9(const float* pooled, int ldp, const float* w_hwio, int Cs, int c_first, int Cp, int D, int B,
                float* tap9 /* [B][9 taps][D]: per-tap vectors, folded into border-class biases by tc_conv3_kernel */,
                cudaStream_t s);
int region_sums_chunks(int HW, int D);
// S[b][tap][n] = sum of g over the output pixels at which `tap` reads inside the image; partial: [B][chunks][9][D]
void region_sums(const float* g, int ldg, int B, int H, int W, int dil, int D, float* partial, float* S, cudaStream_t s);
void pool_wgrad(const float* pooled, int ldp, const float* S, int B, int Cs, int c_first, int Cp, int D, float* dw,
                cudaStream_t s);
void pool_dgrad(const float* S, const float* w_hwio, int Cs, int c_first, int Cp, int D, int B, float scale,
                float* dpooled, int ldo, cudaStream_t s);

// ---------------- misc (k_misc.cu) ----------------
struct ResizeTab { const int32_t* lo; const int32_t* hi; const float* lerp;   // [n_out]
                   const int32_t* g_lo; const int32_t* g_hi; };                // [n_in] gather ranges (bwd)
void bilinear_fwd(const float* x, int ldx, float* y, int ldy, int B, int Hi, int Wi, int Ho, int Wo, int C,
                  ResizeTab ty, ResizeTab tx, cudaStream_t s);
void bilinear_bwd(const float* dy, int lddy, float* dx, int lddx, int B, int Hi, int Wi, int Ho, int Wo, int C,
                  ResizeTab ty, ResizeTab tx, cudaStream_t s);

void head_fwd(const float* x, int ldx, const float* w, const float* bias, const float* drop_mask, float keep_scale,
              float* z, int M, int C, cudaStream_t s);
void head_bwd(const float* x, int ldx, const float* w, const float* drop_mask, float keep_scale, const float* dz,
              float* dx, int lddx, float* partials, float* dw, float* db, int M, int C, cudaStream_t s);

struct LossArgs {
  const float* z_lo;          // [B,h,w,2] low-res logits
  const float* labels;        // pool [n,H,W,2]
  const int32_t* index;       // [B] or null
  int B, h, w, H, W;
  ResizeTab ty, tx;
  int dice; float label_smoothing;
  float* p1;                  // [B,H,W] scratch: softmax prob of channel 1
  float* partials;            // [B][G][4]
  float* coef;                // [B][2] + loss at [2B]
  float* dz_hi;               // [B,H,W,2]
  float* loss_out;            // nullable
  const float* theta; int64_t n_l2; float l2_coef;   // for the reported loss value only
};
void loss_fwd_bwd(const LossArgs& a, cudaStream_t s);

// multi-class head loss with sparse labels (k_mc.cu): class id per example + binary mask, C = n_classes + 1 channels
struct McLossArgs {
  const float* z_lo; int ldz;     // [B*h*w, ldz] low-res logits (columns >= C are padding)
  const float* mask;              // pool [n,H,W] foreground mask (> 0.5 = foreground); nullable for mc_predict
  const int32_t* cls;             // pool [n] class id of the foreground, 1..C-1 (channel 0 = background)
  const int32_t* index;           // [B] or null
  int B, h, w, H, W, C;
  ResizeTab ty, tx;
  int dice; float label_smoothing;
  float* lse; float* pt;          // [B,H,W] scratch: log-sum-exp and target-class probability per pixel
  float* partials;                // [B][chunks][2] (+ 148 L2 partials)
  float* coef;                    // [B]
  float* dz_lo; int lddz;         // out [B*h*w, lddz]
  float* cellgrad;                // scratch [B*h*w][4][lddz]: per-cell corner sums of the backward pass
  float* loss_out;                // nullable
  const float* theta; int64_t n_l2; float l2_coef;
};
int mc_loss_chunks(int H);
void mc_loss_fwd_bwd(const McLossArgs& a, cudaStream_t s);
void mc_predict(const McLossArgs& a, int32_t* class_map, uint32_t* inter, uint32_t* uni, cudaStream_t s);
void mc_pad_head(const float* w, const float* bias, float* wp, float* bp, int K, int C, int Cp, cudaStream_t s);
void mc_unpad_grad(const float* gwp, const float* gbp, float* gw, float* gb, int K, int C, int Cp, cudaStream_t s);
void mc_mul_mask(const float* x, const float* mask, float scale, float* y, int64_t n, cudaStream_t s);
void sumsq_partials(const float* x, int64_t n, float* out, int n_blocks, cudaStream_t s);

void predict_mask_iou(const float* z_lo, const float* labels, const int32_t* index, int B, int h, int w, int H,
                      int W, ResizeTab ty, ResizeTab tx, float* pred_out, float* logits_out, uint32_t* inter,
                      uint32_t* uni, cudaStream_t s);

// optimizer (ApplyAdam beta1=0 / ApplyGradientDescent) over the flat buffer; hyper = {b1p, b2p} device scalars
void scale_buffer(float* x, int64_t n, float s, cudaStream_t st);
void adam_step(float* theta, float* v, const float* g, int64_t n, int64_t n_l2, const float* lr_dev, float* powers,
               float l2_coef, int sgd, cudaStream_t s, float lr_imm = 0.f, float b2p_imm = 0.f);
void delta_accumulate(float* dsum, const float* a, const float* b, int64_t n, int first, cudaStream_t s);
void meta_apply(float* theta, const float* dsum, float scale, int64_t n, cudaStream_t s);
void reduce_partials(const float* partials, int G, int n, float* out, cudaStream_t s);
// out[o * out_stride + i] = sum_g partials[g][o * n_inner + i]   (o < n_outer, i < n_inner)
void reduce_partials_strided(const float* partials, int G, int n_inner, int n_outer, float* out, int64_t out_stride,
                             cudaStream_t s);
void fill_dropout_mask(float* mask, int64_t n, float rate, uint64_t seed, const uint64_t* seed_dev, cudaStream_t s);

}  // namespace mliis
