// Static plan of EfficientLab(efficientnet-b0 truncated at block 10, RSD decoder): variable tables in
// tf.trainable_variables() order, BatchNorm table, activation/gradient workspace layout.
// Host-only (no CUDA); shared by the engine and by table queries on machines without a GPU.
//
// Reference: models/efficientnet/efficientnet_builder.py:90-149, efficientnet_model.py:133-440,
// models/efficientlab.py:23-231.  Expected TF variable names: SURVEY.md section 8a.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace mliis {

struct ParamEntry {
  std::string name;
  int64_t offset;   // into the engine's flat parameter buffer (L2-regularised tensors first, 16B aligned)
  int64_t size;
  int ndim;
  int shape[4];
  int l2;
};

struct BnRef {
  int idx = -1;       // BN layer index (creation order)
  int C = 0;
  int64_t gamma = 0, beta = 0;   // theta offsets
  int off = 0;        // channel offset into the per-channel BN arrays / moving stats
  int fused = 0;
};

struct BnEntry { std::string scope; int C; int off; int fused; };

struct Buf {          // activation buffer in the slot workspace
  int64_t off = -1;   // float offset
  int HW = 0;         // rows per image
  int C = 0;
  int ld = 0;
};

struct BlockPlan {
  int k, stride, cin, cout, ce, cr;
  bool expand, skip;
  float dc_rate;
  int dc_idx;                         // index among drop-connect blocks, -1 if none
  int Hin, Win, Hout, Wout, pad_t, pad_l;
  int64_t w_expand, w_dw, w_se1, b_se1, w_se2, b_se2, w_proj;
  BnRef bn0, bn1, bn2;
  Buf E, D, P, Y;
  int64_t pool, hidpre, gate, dpool;  // [maxB][ce|cr]
  int64_t extra_grad;                 // gradient arriving from the decoder at this block's output (-1: none)
};

struct RsdPlan {
  int r;                 // reduction index (1-based)
  int skip_block;
  int h, w, hin, win;    // output / deep-input spatial size
  int skipC, catC, pyrC, D;
  bool identity_up;
  int64_t w0, b0, w1, b1, w2, b2;
  BnRef bn[3];
  Buf cat, c0, c1, pyr, c2, out;
  int64_t pooled, dpooled;   // [maxB][catC]
  int64_t bias9, S9;         // folded pooled branch: [maxB][9][D] border-class bias / region sums of dL/d(conv2d_2)
  int tab;                   // resize table index (-1 when identity)
};

struct Plan {
  int image_size = 224, maxB = 8;
  int D = 112, n_out = 2;
  std::vector<ParamEntry> params;      // TF creation order
  std::vector<BnEntry> bns;
  std::vector<BlockPlan> blocks;
  std::vector<RsdPlan> rsds;
  std::vector<int> rsd_list;
  int64_t n_params = 0;        // true parameter count (2 071 714)
  int64_t n_theta = 0;         // padded flat length
  int64_t n_l2 = 0;            // first n_l2 floats of theta are L2-regularised
  int n_bn_ch = 0;
  int n_dc = 0;
  // stem
  int64_t w_stem = 0; BnRef bn_stem; Buf S0; int stem_pad_t = 0, stem_pad_l = 0, Hs = 0, Ws = 0;
  // head
  int64_t w_head = 0, b_head = 0;
  int hl = 0, wl = 0;          // low-res logits size
  // workspace (float offsets)
  int64_t z_lo = 0, dz_lo = 0, p1 = 0, dz_hi = 0, dropmask = 0;
  // multi-class head (n_out > 2): logits padded to Cp columns, padded head operand / gradient, lse scratch
  int Cp = 2;
  int64_t mc_lse = 0, mc_wp = 0, mc_bp = 0, mc_gwp = 0, mc_gbp = 0, mc_xdrop = 0, mc_cellgrad = 0;
  int64_t bn_mean = 0, bn_rstd = 0, bn_a = 0, bn_b = 0, bn_k = 0;
  int64_t grads = 0;
  int64_t gY[2] = {0, 0}, gP = 0, gD = 0, gE = 0;
  int64_t g_out = 0, g_c = 0, g_c0 = 0, g_c1 = 0, g_pyr = 0, g_cat = 0, g_up = 0, g_skip = 0, g_deep = 0;
  int64_t partials = 0, partials_len = 0, wT = 0, tn_scratch = 0, dcs = 0, lr_dev = 0, loss_coef = 0;
  // tensor-core operand cache: every dense layer's forward / dgrad operand, rebuilt once per step (tc_prep_all)
  struct PrepJob { int64_t w_off, dst; int taps, Ci, Co, dgrad, decoder, Cs; };   // Cs: source channels per tap
  std::vector<PrepJob> prep_jobs;
  int64_t wcache = 0;
  int64_t ws_floats = 0;
  // resize tables needed: (in, out) pairs
  std::vector<std::pair<int, int>> resize_pairs;
  int tab_final = -1;

  struct Named { std::string name; Buf buf; };
  std::vector<Named> named;   // debug lookup

  void build(int image_size, int max_batch, const int* rsd, float final_dropout_rate, int n_out = 2);
};

void same_pad(int n, int k, int s, int d, int* lo, int* hi);

}  // namespace mliis
