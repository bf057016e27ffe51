#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "kernels.h"

namespace mliis {

void same_pad(int n, int k, int s, int d, int* lo, int* hi) {
  // TF 'SAME' [TF-ext]: total = max((ceil(n/s)-1)*s + (k-1)*d + 1 - n, 0); lo = total/2; hi = total - lo
  int out = (n + s - 1) / s;
  int keff = (k - 1) * d + 1;
  int p = std::max((out - 1) * s + keff - n, 0);
  *lo = p / 2;
  *hi = p - p / 2;
}

namespace {

struct StageDef { int r, k, s, e, i, o; float se; };
// efficientnet_builder.py:130-135
const StageDef kStages[] = {
    {1, 3, 1, 1, 32, 16, 0.25f},  {2, 3, 2, 6, 16, 24, 0.25f},  {2, 5, 2, 6, 24, 40, 0.25f},
    {3, 3, 2, 6, 40, 80, 0.25f},  {3, 5, 1, 6, 80, 112, 0.25f}, {4, 5, 2, 6, 112, 192, 0.25f},
    {1, 3, 1, 6, 192, 320, 0.25f}};
constexpr int kMaxBlockNum = 10;       // models/efficientlab.py:73-75 (efficientnet-b0)
constexpr float kDropConnect = 0.2f;   // efficientnet_builder.py:128

struct Builder {
  Plan& p;
  int64_t ws = 0;
  explicit Builder(Plan& pl) : p(pl) {}

  int add_param(const std::string& name, std::initializer_list<int> shape, int l2) {
    ParamEntry e;
    e.name = name;
    e.ndim = (int)shape.size();
    e.size = 1;
    int i = 0;
    for (int s : shape) { e.shape[i++] = s; e.size *= s; }
    for (; i < 4; ++i) e.shape[i] = 1;
    e.l2 = l2;
    e.offset = -1;
    p.params.push_back(e);
    return (int)p.params.size() - 1;
  }
  // returns index of gamma; beta = +1
  int add_bn(const std::string& scope, int C, int fused, BnRef* ref) {
    int gi = add_param(scope + "/gamma", {C}, 0);
    add_param(scope + "/beta", {C}, 0);
    ref->idx = (int)p.bns.size();
    ref->C = C;
    ref->off = p.n_bn_ch;
    ref->fused = fused;
    ref->gamma = gi;   // temporarily the param index; resolved to offsets later
    ref->beta = gi + 1;
    p.bns.push_back({scope, C, p.n_bn_ch, fused});
    p.n_bn_ch += C;
    return gi;
  }
  int64_t alloc(int64_t n) {
    int64_t o = ws;
    ws += (n + 63) / 64 * 64;   // 256-byte granularity
    return o;
  }
  Buf abuf(const std::string& name, int HW, int C) {
    Buf b;
    b.HW = HW; b.C = C; b.ld = C;
    b.off = alloc((int64_t)p.maxB * HW * C);
    p.named.push_back({name, b});
    return b;
  }
};

}  // namespace

void Plan::build(int image_size_, int max_batch, const int* rsd, float final_dropout_rate, int n_out_) {
  n_out = n_out_;
  image_size = image_size_;
  maxB = max_batch;
  if (image_size % 32 != 0 || image_size < 32) throw std::invalid_argument("image_size must be a positive multiple of 32");
  if (maxB < 1 || maxB > 64) throw std::invalid_argument("max_batch must be in [1, 64]");
  Builder b(*this);
  const std::string pre = "efficientnet-b0/model/";

  // ---- variables in creation order ----
  int pi_stem = b.add_param(pre + "stem/conv2d/kernel", {3, 3, 3, 32}, 1);
  b.add_bn(pre + "stem/tpu_batch_normalization", 32, 0, &bn_stem);

  struct BlkIdx { int we, wd, s1w, s1b, s2w, s2b, wp; };
  std::vector<BlkIdx> bidx;
  int num_blocks = 0;
  int H = image_size / 2;   // after the stem (SAME, stride 2)
  for (const StageDef& st : kStages) {
    num_blocks += st.r;
    if (num_blocks > kMaxBlockNum + 1) break;   // efficientnet_builder.py:105
    for (int rep = 0; rep < st.r; ++rep) {
      BlockPlan bp{};
      bp.k = st.k;
      bp.stride = rep == 0 ? st.s : 1;
      bp.cin = rep == 0 ? st.i : st.o;
      bp.cout = st.o;
      bp.expand = st.e != 1;
      bp.ce = bp.cin * st.e;
      bp.cr = std::max(1, (int)(bp.cin * st.se));   // efficientnet_model.py:203-204
      bp.skip = bp.stride == 1 && bp.cin == bp.cout;
      bp.Hin = bp.Win = H;
      bp.Hout = bp.Wout = (H + bp.stride - 1) / bp.stride;
      int hi;
      same_pad(H, bp.k, bp.stride, 1, &bp.pad_t, &hi);
      bp.pad_l = bp.pad_t;
      H = bp.Hout;
      bp.extra_grad = -1;
      blocks.push_back(bp);
    }
  }
  const int nb = (int)blocks.size();
  for (int i = 0; i < nb; ++i) {
    BlockPlan& bp = blocks[i];
    bp.dc_rate = kDropConnect * (float)i / (float)nb;   // efficientnet_model.py:426-428
    bp.dc_idx = (bp.skip && bp.dc_rate > 0.f) ? n_dc++ : -1;
    const std::string sc = pre + "blocks_" + std::to_string(i) + "/";
    BlkIdx ix{};
    int nconv = 0, nbn = 0;
    auto cname = [&](int n) { return n == 0 ? std::string("conv2d") : "conv2d_" + std::to_string(n); };
    auto bname = [&](int n) {
      return n == 0 ? std::string("tpu_batch_normalization") : "tpu_batch_normalization_" + std::to_string(n);
    };
    ix.we = -1;
    if (bp.expand) {
      ix.we = b.add_param(sc + cname(nconv++) + "/kernel", {1, 1, bp.cin, bp.ce}, 1);
      b.add_bn(sc + bname(nbn++), bp.ce, 0, &bp.bn0);
    }
    ix.wd = b.add_param(sc + "depthwise_conv2d/depthwise_kernel", {bp.k, bp.k, bp.ce, 1}, 1);
    b.add_bn(sc + bname(nbn++), bp.ce, 0, &bp.bn1);
    ix.s1w = b.add_param(sc + "se/conv2d/kernel", {1, 1, bp.ce, bp.cr}, 1);
    ix.s1b = b.add_param(sc + "se/conv2d/bias", {bp.cr}, 1);
    ix.s2w = b.add_param(sc + "se/conv2d_1/kernel", {1, 1, bp.cr, bp.ce}, 1);
    ix.s2b = b.add_param(sc + "se/conv2d_1/bias", {bp.ce}, 1);
    ix.wp = b.add_param(sc + cname(nconv++) + "/kernel", {1, 1, bp.ce, bp.cout}, 1);
    b.add_bn(sc + bname(nbn++), bp.cout, 0, &bp.bn2);
    bidx.push_back(ix);
  }
  // reduction endpoints (efficientnet_model.py:417-439)
  std::vector<int> reduction_block(8, -1);
  {
    int ridx = 0;
    for (int i = 0; i < nb; ++i)
      if (i == nb - 1 || blocks[i + 1].stride > 1) reduction_block[++ridx] = i;
  }
  // decoder (efficientlab.py:153-231): RSD modules in descending reduction order
  for (int i = 0; i < 4 && rsd[i] > 0; ++i) rsd_list.push_back(rsd[i]);
  std::sort(rsd_list.begin(), rsd_list.end(), [](int a, int c) { return a > c; });
  struct RsdIdx { int w0, b0, w1, b1, w2, b2; };
  std::vector<RsdIdx> ridxs;
  int deepC = blocks[reduction_block[4]].cout;
  int deepH = blocks[reduction_block[4]].Hout;
  for (int r : rsd_list) {
    if (r < 1 || r > 4 || reduction_block[r] < 0) throw std::invalid_argument("rsd entries must be in 1..4");
    if (deepC != D) throw std::invalid_argument("deep feature width != 112 (extra 1x1 branch not supported)");
    RsdPlan rp{};
    rp.r = r;
    rp.skip_block = reduction_block[r];
    rp.skipC = blocks[rp.skip_block].cout;
    rp.h = rp.w = blocks[rp.skip_block].Hout;
    rp.hin = rp.win = deepH;
    rp.identity_up = rp.h == rp.hin;
    rp.D = D;
    rp.catC = deepC + rp.skipC;
    rp.pyrC = 2 * D + rp.catC;
    const std::string sc = "decode/decode_skip_connections_" + std::to_string(r - 1) + "/";
    RsdIdx ix{};
    ix.w0 = b.add_param(sc + "conv2d/kernel", {1, 1, rp.catC, D}, 1);
    ix.b0 = b.add_param(sc + "conv2d/bias", {D}, 1);
    b.add_bn(sc + "batch_normalization", D, 1, &rp.bn[0]);
    ix.w1 = b.add_param(sc + "conv2d_1/kernel", {3, 3, rp.catC, D}, 1);
    ix.b1 = b.add_param(sc + "conv2d_1/bias", {D}, 1);
    b.add_bn(sc + "batch_normalization_1", D, 1, &rp.bn[1]);
    ix.w2 = b.add_param(sc + "conv2d_2/kernel", {3, 3, rp.pyrC, D}, 1);
    ix.b2 = b.add_param(sc + "conv2d_2/bias", {D}, 1);
    b.add_bn(sc + "batch_normalization_2", D, 1, &rp.bn[2]);
    ridxs.push_back(ix);
    rsds.push_back(rp);
    deepC = D;
    deepH = rp.h;
  }
  int pi_wh = b.add_param("decode/final_layer_weights/kernel", {1, 1, D, n_out}, 1);
  int pi_bh = b.add_param("decode/final_layer_weights/bias", {n_out}, 1);
  hl = wl = deepH;

  // ---- flat offsets: L2-regularised tensors first, then BN gamma/beta; each 16-byte aligned ----
  int64_t off = 0;
  n_params = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (ParamEntry& e : params) {
      if ((pass == 0) != (e.l2 == 1)) continue;
      e.offset = off;
      off += (e.size + 3) / 4 * 4;
      n_params += e.size;
    }
    if (pass == 0) n_l2 = off;
  }
  n_theta = off;
  auto P = [&](int idx) { return params[idx].offset; };
  auto fix_bn = [&](BnRef& r) {
    if (r.idx < 0) return;
    r.gamma = P((int)r.gamma);
    r.beta = P((int)r.beta);
  };
  w_stem = P(pi_stem);
  fix_bn(bn_stem);
  for (int i = 0; i < nb; ++i) {
    BlockPlan& bp = blocks[i];
    const BlkIdx& ix = bidx[i];
    bp.w_expand = ix.we >= 0 ? P(ix.we) : -1;
    bp.w_dw = P(ix.wd);
    bp.w_se1 = P(ix.s1w); bp.b_se1 = P(ix.s1b); bp.w_se2 = P(ix.s2w); bp.b_se2 = P(ix.s2b);
    bp.w_proj = P(ix.wp);
    fix_bn(bp.bn0); fix_bn(bp.bn1); fix_bn(bp.bn2);
  }
  for (size_t i = 0; i < rsds.size(); ++i) {
    RsdPlan& rp = rsds[i];
    rp.w0 = P(ridxs[i].w0); rp.b0 = P(ridxs[i].b0); rp.w1 = P(ridxs[i].w1); rp.b1 = P(ridxs[i].b1);
    rp.w2 = P(ridxs[i].w2); rp.b2 = P(ridxs[i].b2);
    for (int j = 0; j < 3; ++j) fix_bn(rp.bn[j]);
  }
  w_head = P(pi_wh);
  b_head = P(pi_bh);

  // ---- workspace ----
  const int B = maxB;
  Hs = Ws = image_size / 2;
  { int hi; same_pad(image_size, 3, 2, 1, &stem_pad_t, &hi); stem_pad_l = stem_pad_t; }
  S0 = b.abuf("stem.conv", Hs * Ws, 32);
  int64_t maxY = (int64_t)Hs * Ws * 32, maxD = 0, maxE = 0, maxP = 0;
  int64_t max_partials = 0, max_tn = 0, max_wT = 0;
  auto upd = [](int64_t& m, int64_t v) { if (v > m) m = v; };
  upd(max_partials, (int64_t)rc_num_chunks(B * Hs * Ws, 32) * 2 * 32);
  upd(max_partials, (int64_t)stem_wgrad_blocks(B, Hs, Ws) * 864);
  for (int i = 0; i < nb; ++i) {
    BlockPlan& bp = blocks[i];
    const std::string n = "b" + std::to_string(i);
    const int HWi = bp.Hin * bp.Win, HWo = bp.Hout * bp.Wout;
    if (bp.expand) bp.E = b.abuf(n + ".expand", HWi, bp.ce);
    bp.D = b.abuf(n + ".dw", HWo, bp.ce);
    bp.P = b.abuf(n + ".project", HWo, bp.cout);
    bp.Y = b.abuf(n + ".out", HWo, bp.cout);
    bp.pool = b.alloc((int64_t)B * bp.ce);
    bp.hidpre = b.alloc((int64_t)B * bp.cr);
    bp.gate = b.alloc((int64_t)B * bp.ce);
    bp.dpool = b.alloc((int64_t)B * bp.ce);
    { Buf g; g.off = bp.gate; g.HW = 1; g.C = bp.ce; g.ld = bp.ce; named.push_back({n + ".gate", g}); }
    upd(maxY, (int64_t)HWo * bp.cout);
    upd(maxY, (int64_t)HWi * bp.cin);
    upd(maxP, (int64_t)HWo * bp.cout);
    upd(maxD, (int64_t)HWo * bp.ce);
    upd(maxE, (int64_t)HWi * bp.ce);
    // scratch needs
    upd(max_partials, (int64_t)rc_num_chunks(B * HWi, bp.ce) * 2 * bp.ce);
    upd(max_partials, (int64_t)rc_num_chunks(B * HWo, bp.ce) * 2 * bp.ce);
    upd(max_partials, (int64_t)rc_num_chunks(B * HWo, bp.cout) * 2 * bp.cout);
    upd(max_partials, (int64_t)B * rc_num_img_chunks(HWo, bp.ce) * bp.ce + (int64_t)B * (bp.ce + bp.cr) + 64);
    upd(max_partials, (int64_t)dw_wgrad_blocks(B, bp.Hout, bp.Wout, bp.stride) * bp.k * bp.k * bp.ce);
    if (bp.expand) {
      upd(max_tn, (int64_t)gemm_tn_scratch(B * HWi, bp.cin, bp.ce, 0));
      upd(max_tn, (int64_t)tc_wgrad_scratch(0, B * HWi, B, bp.Hin, bp.Win, bp.ce, bp.cin, 1));   // swapped roles
      upd(max_wT, (int64_t)bp.cin * bp.ce);
    }
    upd(max_tn, (int64_t)gemm_tn_scratch(B * HWo, bp.ce, bp.cout, 0));
    upd(max_tn, (int64_t)tc_wgrad_scratch(0, B * HWo, B, bp.Hout, bp.Wout, bp.ce, bp.cout, 1));
    upd(max_wT, (int64_t)bp.ce * bp.cout);
  }
  int64_t maxDec = 0, maxPyr = 0, maxCat = 0, maxSkip = 0, maxDeep = 0;
  for (RsdPlan& rp : rsds) {
    const std::string n = "rsd" + std::to_string(rp.r - 1);
    const int HW = rp.h * rp.w;
    rp.cat = b.abuf(n + ".cat", HW, rp.catC);
    rp.c0 = b.abuf(n + ".conv2d", HW, D);
    rp.c1 = b.abuf(n + ".conv2d_1", HW, D);
    rp.pyr = b.abuf(n + ".pyr", HW, rp.pyrC);
    rp.c2 = b.abuf(n + ".conv2d_2", HW, D);
    rp.out = b.abuf(n + ".out", HW, D);
    { Buf u = rp.cat; u.C = D; named.push_back({n + ".up", u}); }
    rp.pooled = b.alloc((int64_t)B * rp.catC);
    rp.dpooled = b.alloc((int64_t)B * rp.catC);
    rp.bias9 = b.alloc((int64_t)B * 9 * D);
    rp.S9 = b.alloc((int64_t)B * 9 * D);
    upd(max_partials, (int64_t)B * region_sums_chunks(HW, D) * 9 * D);
    rp.tab = -1;
    if (!rp.identity_up) {
      rp.tab = (int)resize_pairs.size();
      resize_pairs.push_back({rp.hin, rp.h});
    }
    upd(maxDec, (int64_t)HW * D);
    upd(maxPyr, (int64_t)HW * rp.pyrC);
    upd(maxCat, (int64_t)HW * rp.catC);
    upd(maxSkip, (int64_t)HW * rp.skipC);
    upd(maxDeep, (int64_t)rp.hin * rp.win * D);
    upd(max_partials, (int64_t)rc_num_chunks(B * HW, D) * 2 * D);
    upd(max_partials, (int64_t)B * rc_num_img_chunks(HW, rp.catC) * rp.catC);
    upd(max_tn, (int64_t)gemm_tn_scratch(B * HW, rp.catC, D, 0));
    upd(max_tn, (int64_t)gemm_tn_scratch(B * HW, 9 * rp.catC, D, 1));
    upd(max_tn, (int64_t)gemm_tn_scratch(B * HW, 9 * rp.pyrC, D, 1));
    upd(max_tn, (int64_t)tc_wgrad_scratch(0, B * HW, B, rp.h, rp.w, rp.catC, D, 1));
    upd(max_tn, (int64_t)tc_wgrad_scratch(1, B * HW, B, rp.h, rp.w, rp.catC, D, 9));
    upd(max_tn, (int64_t)tc_wgrad_scratch(1, B * HW, B, rp.h, rp.w, rp.pyrC, D, 9));
    upd(max_wT, (int64_t)9 * rp.pyrC * D);
    // the decoder's gradient wrt a backbone block output
    BlockPlan& sb = blocks[rp.skip_block];
    if (sb.extra_grad < 0) sb.extra_grad = b.alloc((int64_t)B * HW * sb.cout);
  }
  if (rsds.empty()) throw std::invalid_argument("at least one --rsd entry is required");
  {
    BlockPlan& last = blocks[nb - 1];
    if (last.extra_grad < 0) last.extra_grad = b.alloc((int64_t)B * last.Hout * last.Wout * last.cout);
  }
  tab_final = (int)resize_pairs.size();
  resize_pairs.push_back({hl, image_size});
  const int HWl = hl * wl, HWf = image_size * image_size;
  Cp = n_out == 2 ? 2 : (n_out + 3) / 4 * 4;
  z_lo = b.alloc((int64_t)B * HWl * Cp);
  { Buf z; z.off = z_lo; z.HW = HWl; z.C = n_out; z.ld = Cp; named.push_back({"head.logits_lowres", z}); }
  dz_lo = b.alloc((int64_t)B * HWl * Cp);
  { Buf z; z.off = dz_lo; z.HW = HWl; z.C = n_out; z.ld = Cp; named.push_back({"head.dlogits_lowres", z}); }
  p1 = b.alloc((int64_t)B * HWf);
  if (n_out == 2) {
    dz_hi = b.alloc((int64_t)B * HWf * 2);
  } else {
    mc_lse = b.alloc((int64_t)B * HWf);
    mc_cellgrad = b.alloc((int64_t)B * HWl * 4 * Cp);
    mc_wp = b.alloc((int64_t)D * Cp);
    mc_bp = b.alloc(Cp);
    mc_gwp = b.alloc((int64_t)D * Cp);
    mc_gbp = b.alloc(Cp);
    if (final_dropout_rate > 0.f) mc_xdrop = b.alloc((int64_t)B * HWl * D);
    upd(max_wT, (int64_t)D * Cp);
    upd(max_tn, (int64_t)gemm_tn_scratch(B * HWl, D, Cp, 0));
    upd(max_tn, (int64_t)tc_wgrad_scratch(0, B * HWl, B, hl, wl, Cp, D, 1));      // swapped roles
    upd(max_partials, (int64_t)B * HWl * 2 + 148 + 64);      // one (CE, I) pair per low-res cell + L2 partials
    upd(max_partials, (int64_t)rc_num_img_chunks(B * HWl, Cp) * Cp + 64);
  }
  dropmask = final_dropout_rate > 0.f ? b.alloc((int64_t)B * HWl * D) : -1;
  if (dropmask >= 0) { Buf m; m.off = dropmask; m.HW = HWl; m.C = D; m.ld = D; named.push_back({"head.dropmask", m}); }
  bn_mean = b.alloc(n_bn_ch);
  bn_rstd = b.alloc(n_bn_ch);
  bn_a = b.alloc(n_bn_ch);
  bn_b = b.alloc(n_bn_ch);
  bn_k = b.alloc(2 * 1024);
  grads = b.alloc(n_theta);
  gY[0] = b.alloc(B * maxY);
  gY[1] = b.alloc(B * maxY);
  gP = b.alloc(B * maxP);
  gD = b.alloc(B * maxD);
  gE = b.alloc(B * maxE);
  g_out = b.alloc(B * maxDec);
  g_c = b.alloc(B * maxDec);
  g_c0 = b.alloc(B * maxDec);
  g_c1 = b.alloc(B * maxDec);
  g_up = b.alloc(B * maxDec);
  g_pyr = b.alloc(B * maxPyr);
  g_cat = b.alloc(B * maxCat);
  g_skip = b.alloc(B * std::max(maxSkip, maxDec));
  g_deep = b.alloc(B * maxDeep);
  // loss partials, head partials (+ staging row)
  upd(max_partials, (int64_t)B * 32 * 4 + 148 + 64);
  upd(max_partials, (int64_t)297 * (D * 2 + 2) + 64);
  partials_len = max_partials + 1024;
  partials = b.alloc(partials_len);
  wT = b.alloc(2 * max_wT);   // hi (+ lo) operand planes of the tensor-core path
  tn_scratch = b.alloc(max_tn);
  {
    int64_t total = 0;
    auto add = [&](int64_t w_off, int taps, int Ci, int Co, int decoder, int Cs = 0) {
      for (int dgrad = 0; dgrad < 2; ++dgrad) {
        prep_jobs.push_back({w_off, total, taps, Ci, Co, dgrad, decoder, Cs > 0 ? Cs : Ci});
        total += ((int64_t)2 * taps * Ci * Co + 31) / 32 * 32;
      }
    };
    for (const BlockPlan& bp : blocks) {
      if (bp.expand) add(bp.w_expand, 1, bp.cin, bp.ce, 0);
      add(bp.w_proj, 1, bp.ce, bp.cout, 0);
    }
    for (const RsdPlan& rp : rsds) {
      add(rp.w0, 1, rp.catC, D, 1);
      add(rp.w1, 9, rp.catC, D, 1);
      add(rp.w2, 9, 2 * D, D, 1, rp.pyrC);   // conv2d_2 WITHOUT the pooled channels (folded: k_pool.cu)
    }
    wcache = b.alloc(total);
    for (PrepJob& j : prep_jobs) j.dst += 0;   // offsets are relative to wcache
  }
  dcs = b.alloc((int64_t)std::max(1, n_dc) * B);
  lr_dev = b.alloc(64);
  loss_coef = b.alloc(2 * B + 64);
  ws_floats = b.ws;
}

}  // namespace mliis
