"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): logits max-abs <= 1e-2, adapted weights rel-L2 <= 1e-3, integer work
(masks given logits, IoU counts) bit-exact.  The fp32 engine is held to much tighter bounds than those.
The oracle runs in float64; the engine in fp32.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.efficientlab_oracle import (Arch, EfficientLabOracle, OptState, conv2d_same, depthwise_same,
                                        iou_counts, resize_bilinear_ac, swish)
from tests.parity_util import make_engine, make_problem, per_param_report, rel_err, rel_l2, split_vars

pytestmark = pytest.mark.gpu


def _dev(x):
    return torch.as_tensor(x, dtype=torch.float32).cuda().contiguous()


# ------------------------------------------------------------------------------------------------
# single kernels through their C-ABI entry points
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,stride,H,C", [(3, 1, 28, 32), (3, 2, 28, 96), (5, 1, 14, 48), (5, 2, 56, 144),
                                          (3, 1, 9, 16), (5, 2, 11, 40)])
def test_dwconv_fwd(k, stride, H, C):
    import ctypes
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(k * 100 + stride * 10 + H)
    B = 2
    x = torch.randn(B, H, H, C, generator=g)
    w = torch.randn(k, k, C, 1, generator=g) * 0.3
    a = torch.rand(C, generator=g) + 0.5
    b = torch.randn(C, generator=g) * 0.2
    ref = depthwise_same(swish(x.double() * a.double() + b.double()).permute(0, 3, 1, 2), w.double(), stride)
    ref = ref.permute(0, 2, 3, 1)
    Ho = (H + stride - 1) // stride
    xd, wd, ad, bd = _dev(x), _dev(w), _dev(a), _dev(b)
    y = torch.empty(B, Ho, Ho, C, device="cuda")
    N.check(N.lib().mliis_dwconv_fwd(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), B, H, H, C, k, stride,
                                     ad.data_ptr(), bd.data_ptr(), None))
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 2e-5
    # without the BN+swish prologue
    ref2 = depthwise_same(x.double().permute(0, 3, 1, 2), w.double(), stride).permute(0, 2, 3, 1)
    N.check(N.lib().mliis_dwconv_fwd(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), B, H, H, C, k, stride, None, None,
                                     None))
    torch.cuda.synchronize()
    assert rel_err(y, ref2) < 2e-5


@pytest.mark.parametrize("M,K,N_", [(1568, 112, 672), (300, 16, 96), (1000, 672, 112), (77, 24, 144)])
def test_gemm_nn(M, K, N_):
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(K, N_, generator=g)
    ref = a.double() @ w.double()
    ad, wd = _dev(a), _dev(w)
    c = torch.empty(M, N_, device="cuda")
    N.check(N.lib().mliis_gemm_nn(ad.data_ptr(), wd.data_ptr(), c.data_ptr(), M, K, N_, 0, None))
    torch.cuda.synchronize()
    assert rel_err(c, ref) < 1e-5


@pytest.mark.parametrize("H,Cin,Cout,dil", [(14, 224, 112, 2), (14, 448, 112, 1), (20, 136, 112, 2), (9, 8, 12, 1)])
def test_conv3x3_fwd(H, Cin, Cout, dil):
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(H + Cin)
    B = 2
    x = torch.randn(B, H, H, Cin, generator=g)
    w = torch.randn(3, 3, Cin, Cout, generator=g) * 0.05
    bias = torch.randn(Cout, generator=g)
    ref = conv2d_same(x.double().permute(0, 3, 1, 2), w.double(), dilation=dil, bias=bias.double()).permute(0, 2, 3, 1)
    xd, wd, bd = _dev(x), _dev(w), _dev(bias)
    y = torch.empty(B, H, H, Cout, device="cuda")
    N.check(N.lib().mliis_conv3x3_fwd(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout,
                                      dil, 0, None))
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-5


def test_bilinear_fwd():
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 14, 14, 112, generator=g)
    ref = resize_bilinear_ac(x.double().permute(0, 3, 1, 2), 56, 56).permute(0, 2, 3, 1)
    # independent cross-check of the oracle itself
    ref_t = F.interpolate(x.double().permute(0, 3, 1, 2), size=(56, 56), mode="bilinear", align_corners=True)
    assert rel_err(ref, ref_t.permute(0, 2, 3, 1)) < 1e-6
    xd = _dev(x)
    y = torch.empty(2, 56, 56, 112, device="cuda")
    N.check(N.lib().mliis_bilinear_fwd(xd.data_ptr(), y.data_ptr(), 2, 14, 14, 56, 56, 112, None))
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-6


def test_adam_step_kernel():
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(5)
    n, n_l2 = 10001, 6000
    th = torch.randn(n, generator=g)
    v = torch.rand(n, generator=g) * 1e-3
    gr = torch.randn(n, generator=g) * 1e-2
    lr, b2p, l2 = 1e-3, 0.999 ** 3, 0.0005
    st = OptState(n, torch.float64)
    st.v = v.double().clone()
    st.b2p = b2p
    mask = torch.zeros(n, dtype=torch.float64)
    mask[:n_l2] = 1
    ref = st.apply(th.double(), gr.double() + l2 * mask * th.double(), lr)
    thd, vd, gd = _dev(th), _dev(v), _dev(gr)
    N.check(N.lib().mliis_adam_step(thd.data_ptr(), vd.data_ptr(), gd.data_ptr(), n, n_l2, lr, b2p, l2, None))
    torch.cuda.synchronize()
    assert rel_err(thd, ref) < 1e-6
    assert rel_err(vd, st.v) < 1e-6


# ------------------------------------------------------------------------------------------------
# whole network
# ------------------------------------------------------------------------------------------------
def _forward_report(size, B, training, dc=None):
    arch, theta, bn, images, labels = make_problem(size, B)
    orc = EfficientLabOracle(arch, torch.float64)
    taps = {}
    dc_t = None
    if dc is not None:
        dc_t = torch.tensor(dc, dtype=torch.float64)
    logits_ref, bn_ref = orc.forward(theta, bn, torch.from_numpy(images), training, dc_masks=dc_t, taps=taps)
    eng = make_engine(arch, theta, bn, size, B)
    xd = _dev(images)
    dcd = _dev(np.asarray(dc, np.float32).reshape(-1)) if dc is not None else None
    logits = eng.forward(0, xd, training, dc_mask=dcd)
    torch.cuda.synchronize()
    rows = []
    for name, ref in taps.items():
        try:
            got = eng.debug_buffer(0, name, B)
        except Exception:
            continue
        refv = ref.reshape(B, -1, ref.shape[-1])
        rows.append((name, rel_err(got, refv)))
    return arch, eng, rows, logits, logits_ref, bn_ref


@pytest.mark.parametrize("size,B,training", [(64, 2, True), (64, 3, False), (96, 2, True)])
def test_forward_layers(size, B, training):
    arch, eng, rows, logits, logits_ref, bn_ref = _forward_report(size, B, training)
    bad = [(n, e) for n, e in rows if not e < 2e-4]
    print("\n".join("%-24s %.3e" % r for r in rows))
    assert not bad, "first mismatching layers: %s" % bad[:4]
    assert (logits.cpu().double() - logits_ref).abs().max().item() < 1e-3
    if training:
        assert rel_err(eng.bn_state(0).cpu(), bn_ref) < 1e-5


def test_forward_drop_connect_masks():
    dc = [[1, 0], [0, 1], [1, 1], [0, 0], [1, 0], [1, 1]]
    arch, eng, rows, logits, logits_ref, _ = _forward_report(64, 2, True, dc=dc)
    bad = [(n, e) for n, e in rows if not e < 2e-4]
    assert not bad, bad[:4]
    assert (logits.cpu().double() - logits_ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("size,B,dice,l2", [(64, 2, True, True), (64, 3, False, False), (96, 2, True, True)])
def test_loss_and_gradients(size, B, dice, l2):
    arch, theta, bn, images, labels = make_problem(size, B)
    orc = EfficientLabOracle(arch, torch.float64, dice=dice, l2=l2)
    loss_ref, g_ref, bn_ref, _ = orc.loss_and_grad(theta, bn, torch.from_numpy(images), torch.from_numpy(labels))
    if l2:   # the engine folds the L2 gradient into the optimizer: compare the data-term gradient
        g_ref = g_ref - 0.0005 * arch.l2_mask() * theta
    eng = make_engine(arch, theta, bn, size, B, dice=dice, l2=l2)
    xd, yd = _dev(images), _dev(labels)
    eng.forward(0, xd, True, want_logits=False)
    loss, grads = eng.loss_backward(0, yd, B)
    torch.cuda.synchronize()
    g = eng.tf_order_vector(grads).cpu().double()
    rep = per_param_report(arch, g, g_ref)
    print("\n".join("%.3e  |g|max=%.3e  %s" % r for r in rep))
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    assert rel_l2(g, g_ref) < 1e-4
    assert all(r[0] < 2e-3 or r[1] < 1e-9 for r in rep), rep[:4]


@pytest.mark.parametrize("sgd,steps", [(False, 3), (True, 3)])
def test_train_steps_weights(sgd, steps):
    size, B = 64, 4
    arch, theta, bn, images, labels = make_problem(size, 8)
    orc = EfficientLabOracle(arch, torch.float64)
    opt = OptState(arch.n_params, torch.float64, sgd=sgd)
    eng = make_engine(arch, theta, bn, size, B, sgd=sgd)
    xd, yd = _dev(images), _dev(labels)
    rng = np.random.default_rng(0)
    th, bns = theta, bn
    lr = 1e-3
    for s in range(steps):
        idx = rng.permutation(8)[:B].astype(np.int32)
        loss, g, bns, _ = orc.loss_and_grad(th, bns, torch.from_numpy(images[idx]), torch.from_numpy(labels[idx]))
        th = opt.apply(th, g, lr)
        eng.train_step(0, xd, yd, lr, index=torch.from_numpy(idx).cuda())
    torch.cuda.synchronize()
    got = eng.tf_order_vector(eng.theta(0)).cpu().double()
    # north_star: adapted weights within 1e-3 relative L2
    assert rel_l2(got, th) < 1e-3
    assert rel_l2(got - theta, th - theta) < 5e-2      # the UPDATE itself agrees, not just theta
    assert rel_err(eng.bn_state(0).cpu(), bns) < 1e-4
    if not sgd:
        assert abs(eng.powers(0)[1].item() - 0.999 ** (steps + 1)) < 1e-6


def test_predict_mask_and_iou_counts():
    size, B = 64, 5
    arch, theta, bn, images, labels = make_problem(size, B)
    orc = EfficientLabOracle(arch, torch.float64)
    # a couple of oracle steps so that moving statistics and logits are not degenerate
    opt = OptState(arch.n_params, torch.float64)
    th, bns = theta, bn
    for _ in range(2):
        _, g, bns, _ = orc.loss_and_grad(th, bns, torch.from_numpy(images), torch.from_numpy(labels))
        th = opt.apply(th, g, 1e-3)
    th = th.float().double()
    bns = bns.float().double()
    pred_ref, logits_ref = orc.predict(th, bns, torch.from_numpy(images))
    eng = make_engine(arch, th, bns, size, B)
    xd, yd = _dev(images), _dev(labels)
    pred, logits, inter, uni = eng.predict(0, xd, yd, want_logits=True)
    torch.cuda.synchronize()
    assert (logits.cpu().double() - logits_ref).abs().max().item() < 1e-2
    # integer work is bit-exact GIVEN the same logits: recompute the reference mask from the engine's logits
    lg = logits.cpu()
    probs = torch.softmax(lg, dim=-1)
    mask_from_engine_logits = (probs > 0.5).float()
    assert torch.equal(pred.cpu(), mask_from_engine_logits)
    for j in range(B):
        i_ref, u_ref = iou_counts(pred.cpu().numpy()[j], labels[j])
        assert int(inter[j].item()) == i_ref and int(uni[j].item()) == u_ref
    # end-to-end disagreement with the float64 oracle mask, reported as a pixel count
    diff = int((pred.cpu() != pred_ref).sum().item())
    assert diff <= 0.001 * pred_ref.numel(), "mask disagreement %d px" % diff


def test_adapt_eval_task_matches_stepwise():
    size, B, T = 64, 4, 3
    arch, theta, bn, images, labels = make_problem(size, 10)
    eng = make_engine(arch, theta, bn, size, max(B, 5))
    xd, yd = _dev(images), _dev(labels)
    init = eng.states[0].clone()
    rng = np.random.default_rng(1)
    bidx = np.stack([rng.permutation(5)[:B] for _ in range(T)]).astype(np.int32)
    qidx = np.arange(5, 10, dtype=np.int32)
    lrs = torch.full((T,), 1e-3, device="cuda")
    inter = torch.zeros(5, dtype=torch.int32, device="cuda")
    uni = torch.zeros(5, dtype=torch.int32, device="cuda")
    eng.adapt_eval_task(0, init, xd, yd, torch.from_numpy(bidx.reshape(-1)).cuda(), lrs, T, B,
                        torch.from_numpy(qidx).cuda(), inter, uni)
    torch.cuda.synchronize()
    th_task = eng.theta(0).clone()
    # the same thing step by step through the session-level entry points
    eng.states[0].copy_(init)
    for t in range(T):
        eng.train_step(0, xd, yd, 1e-3, index=torch.from_numpy(bidx[t]).cuda())
    _, _, inter2, uni2 = eng.predict(0, xd, yd, index=torch.from_numpy(qidx).cuda(), want_pred=False)
    torch.cuda.synchronize()
    assert torch.equal(th_task, eng.theta(0))
    assert torch.equal(inter, inter2) and torch.equal(uni, uni2)


def test_meta_update_kernels():
    arch, theta, bn, images, labels = make_problem(64, 2)
    eng = make_engine(arch, theta, bn, 64, 2)
    n = eng.n_theta
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(n, device="cuda", generator=g)
    b = torch.randn(n, device="cuda", generator=g)
    c = torch.randn(n, device="cuda", generator=g)
    dsum = torch.empty(n, device="cuda")
    eng.delta_accumulate(dsum, a, b, True)
    eng.delta_accumulate(dsum, c, b, False)
    th = torch.randn(n, device="cuda", generator=g)
    th0 = th.clone()
    eng.meta_apply(th, dsum, 0.05)
    torch.cuda.synchronize()
    ref = th0.double() + 0.05 * ((a.double() - b.double()) + (c.double() - b.double()))
    assert rel_err(th, ref) < 1e-6


def test_full_size_one_step():
    """Canonical shape: 224x224, B=8 (one Adam step; the float64 oracle takes a few seconds)."""
    size, B = 224, 8
    arch, theta, bn, images, labels = make_problem(size, B)
    orc = EfficientLabOracle(arch, torch.float64)
    loss_ref, g_ref, bn_ref, logits_ref = orc.loss_and_grad(theta, bn, torch.from_numpy(images), torch.from_numpy(labels))
    g_data = g_ref - 0.0005 * arch.l2_mask() * theta
    eng = make_engine(arch, theta, bn, size, B)
    xd, yd = _dev(images), _dev(labels)
    logits = eng.forward(0, xd, True)
    loss, grads = eng.loss_backward(0, yd, B)
    torch.cuda.synchronize()
    assert (logits.cpu().double() - logits_ref).abs().max().item() < 1e-2
    g = eng.tf_order_vector(grads).cpu().double()
    assert rel_l2(g, g_data) < 1e-4
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))


# ------------------------------------------------------------------------------------------------
# tcgen05 / TMA / TMEM path (kind::tf32): the hardware reads the top 19 bits of every fp32 operand
# ------------------------------------------------------------------------------------------------
def _tf32_rn(x):
    """Round to nearest TF32, ties away from zero (what tc_prep_weights / the transform warps produce)."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _tf32_trunc(x):
    """What tcgen05.mma kind::tf32 does to a raw fp32 operand: the 13 low mantissa bits are ignored.  Single-pass
    mode feeds activations straight from TMA to the MMA (no transform stage); weights are pre-rounded (RN)."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,K,N_", [(1568, 224, 112), (25088, 136, 112), (300, 112, 360), (128, 32, 16), (77, 8, 24)])
def test_tc_gemm_nn(M, K, N_):
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(K, N_, generator=g)
    ref_t = _tf32_trunc(a).double() @ _tf32_rn(w).double()
    ref = a.double() @ w.double()
    ad, wd = _dev(a), _dev(w)
    c = torch.full((M, N_), float("nan"), device="cuda")
    N.check(N.lib().mliis_gemm_nn(ad.data_ptr(), wd.data_ptr(), c.data_ptr(), M, K, N_, 1, None))
    torch.cuda.synchronize()
    assert rel_err(c, ref_t) < 2e-5, "layout / descriptor error"
    assert rel_err(c, ref) < 3e-3
    c.fill_(float("nan"))
    N.check(N.lib().mliis_gemm_nn(ad.data_ptr(), wd.data_ptr(), c.data_ptr(), M, K, N_, 2, None))   # 3xTF32
    torch.cuda.synchronize()
    assert rel_err(c, ref) < 1e-4


@pytest.mark.parametrize("H,Cin,Cout,dil,B", [(56, 136, 112, 2, 2), (56, 360, 112, 1, 1), (14, 224, 112, 2, 3),
                                              (14, 448, 112, 1, 2), (56, 112, 360, 1, 1), (14, 112, 448, 2, 2),
                                              (20, 40, 16, 1, 2), (80, 16, 16, 1, 1)])
def test_tc_conv3x3(H, Cin, Cout, dil, B):
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(H + Cin + Cout)
    x = torch.randn(B, H, H, Cin, generator=g)
    w = torch.randn(3, 3, Cin, Cout, generator=g) * 0.05
    bias = torch.randn(Cout, generator=g)
    ref_t = conv2d_same(_tf32_trunc(x).double().permute(0, 3, 1, 2), _tf32_rn(w).double(), dilation=dil,
                        bias=bias.double()).permute(0, 2, 3, 1)
    xd, wd, bd = _dev(x), _dev(w), _dev(bias)
    y = torch.full((B, H, H, Cout), float("nan"), device="cuda")
    N.check(N.lib().mliis_conv3x3_fwd(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout,
                                      dil, 1, None))
    torch.cuda.synchronize()
    assert rel_err(y, ref_t) < 2e-5, "layout / descriptor / padding error"
    ref = conv2d_same(x.double().permute(0, 3, 1, 2), w.double(), dilation=dil, bias=bias.double()).permute(0, 2, 3, 1)
    y.fill_(float("nan"))
    N.check(N.lib().mliis_conv3x3_fwd(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout,
                                      dil, 2, None))                                                   # 3xTF32
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-4      # fp32 accumulation over 3*9*Cin terms in the tensor-core adder


@pytest.mark.parametrize("H,Cin,Cout,taps,dil,B", [(56, 136, 112, 9, 2, 2), (56, 360, 112, 9, 1, 1), (14, 224, 112, 9, 2, 3),
                                                   (14, 448, 112, 9, 1, 2), (56, 136, 112, 1, 1, 2), (28, 240, 40, 1, 1, 2),
                                                   (112, 96, 16, 1, 1, 1), (14, 672, 112, 1, 1, 8), (20, 40, 16, 9, 1, 2),
                                                   # compact stages: 1-3 channel groups of A, two CTAs per SM
                                                   (56, 16, 96, 1, 1, 2), (56, 32, 16, 1, 1, 2), (28, 72, 24, 1, 1, 3),
                                                   (28, 40, 240, 1, 1, 2)])
def test_tc_wgrad(H, Cin, Cout, taps, dil, B):
    from mliis_b200 import native as N
    g = torch.Generator().manual_seed(H + Cin + Cout + taps)
    a = torch.randn(B, H, H, Cin, generator=g)
    gr = torch.randn(B, H, H, Cout, generator=g)
    ad, gd = _dev(a), _dev(gr)
    # reference: autograd of the SAME-padded conv w.r.t. the weights
    w = torch.zeros(3 if taps == 9 else 1, 3 if taps == 9 else 1, Cin, Cout, dtype=torch.float64, requires_grad=True)
    y = conv2d_same(a.double().permute(0, 3, 1, 2), w, dilation=dil)
    (ref,) = torch.autograd.grad(y, w, gr.double().permute(0, 3, 1, 2))
    ref = ref.reshape(taps * Cin, Cout)
    for mode, tol in ((2, 1e-4), (1, 5e-3)):
        dw = torch.full((taps * Cin, Cout), float("nan"), device="cuda")
        N.check(N.lib().mliis_tc_wgrad(ad.data_ptr(), gd.data_ptr(), dw.data_ptr(), B, H, H, Cin, Cout, taps, dil, mode, None))
        torch.cuda.synchronize()
        assert rel_err(dw, ref) < tol, (mode, rel_err(dw, ref))


def _tf32_network_errors(size, B, steps, mode, warm=0):
    from mliis_b200 import native as N
    arch, theta, bn, images, labels = make_problem(size, 10)
    orc = EfficientLabOracle(arch, torch.float64)
    if warm:   # a representative state: `warm` oracle steps from the random init (SURVEY.md 8d "checkpoint")
        w_opt = OptState(arch.n_params, torch.float64)
        wx, wy = torch.from_numpy(images[:B]), torch.from_numpy(labels[:B])
        for _ in range(warm):
            _, g, bn, _ = orc.loss_and_grad(theta, bn, wx, wy)
            theta = w_opt.apply(theta, g, 1e-3)
        theta, bn = theta.float().double(), bn.float().double()
    opt = OptState(arch.n_params, torch.float64)
    eng = make_engine(arch, theta, bn, size, max(B, 5), gemm_mode=mode)
    xd, yd = _dev(images), _dev(labels)
    # forward + gradient of the first batch
    idx0 = np.arange(B, dtype=np.int32)
    loss_ref, g_ref, _, logits_ref = orc.loss_and_grad(theta, bn, torch.from_numpy(images[idx0]), torch.from_numpy(labels[idx0]))
    g_data = g_ref - 0.0005 * arch.l2_mask() * theta
    logits = eng.forward(0, xd, True, index=torch.from_numpy(idx0).cuda())
    loss, grads = eng.loss_backward(0, yd, B, index=torch.from_numpy(idx0).cuda())
    torch.cuda.synchronize()
    e_logits = (logits.cpu().double() - logits_ref).abs().max().item()
    e_grad = rel_l2(eng.tf_order_vector(grads).cpu().double(), g_data)
    # adapted weights after `steps` Adam steps
    eng.init_state(0, split_vars(arch, theta), bn[0].numpy(), bn[1].numpy())
    rng = np.random.default_rng(0)
    th, bns = theta, bn
    for s in range(steps):
        idx = rng.permutation(5)[:B].astype(np.int32) if B <= 5 else rng.integers(0, 5, B).astype(np.int32)
        _, g, bns, _ = orc.loss_and_grad(th, bns, torch.from_numpy(images[idx]), torch.from_numpy(labels[idx]))
        th = opt.apply(th, g, 1e-3)
        eng.train_step(0, xd, yd, 1e-3, index=torch.from_numpy(idx).cuda())
    q = torch.arange(5, 10, dtype=torch.int32).cuda()
    pred, lg, inter, uni = eng.predict(0, xd, yd, index=q, want_logits=True)
    torch.cuda.synchronize()
    pred_ref, lg_ref = orc.predict(th.float().double(), bns.float().double(), torch.from_numpy(images[5:10]))
    e_theta = rel_l2(eng.tf_order_vector(eng.theta(0)).cpu().double(), th)
    e_pred_logits = (lg.cpu().double() - lg_ref).abs().max().item()
    ious_e, ious_r = [], []
    for j in range(5):
        ious_e.append((inter[j].item() + 1e-7) / (uni[j].item() + 1e-7))
        i_r, u_r = iou_counts(pred_ref.numpy()[j], labels[5 + j])
        ious_r.append((i_r + 1e-7) / (u_r + 1e-7))
    return dict(logits=e_logits, grad=e_grad, theta=e_theta, pred_logits=e_pred_logits,
                miou_engine=float(np.mean(ious_e)), miou_oracle=float(np.mean(ious_r)))


@pytest.mark.parametrize("size,B,mode,warm", [(64, 4, 2, 0), (224, 8, 2, 0), (64, 4, 1, 0), (224, 8, 1, 6)])
def test_tensor_core_modes_within_north_star_tolerances(size, B, mode, warm):
    """3xTF32 (mode 2) must meet every north-star tolerance literally; plain TF32 (mode 1) must meet the weight
    tolerance - its logits error at random-init logit scale (|z| ~ 30) is reported, see DESIGN.md."""
    e = _tf32_network_errors(size, B, 5, mode, warm)
    print("gemm_mode %d, %dx%d B=%d warm=%d: %s" % (mode, size, size, B, warm, e))
    try:
        with open("gpurun_out/tf32_parity.log", "a") as f:
            f.write("gemm_mode %d, %dx%d B=%d warm=%d: %s\n" % (mode, size, size, B, warm, e))
    except OSError:
        pass
    if mode == 2:
        assert e["theta"] < 1e-3             # adapted weights rel-L2 after 5 inner Adam steps
        assert e["grad"] < 1e-3
        assert abs(e["miou_engine"] - e["miou_oracle"]) < 0.005     # per-task mIoU within 0.5 points
        assert e["logits"] < 1e-2        # logits max-abs
    else:
        # MLIIS_GEMM_TF32 = single-pass TF32 on the decoder convs (3xTF32 on the backbone): meets the weight / mIoU
        # bounds; its same-weights logits error (0.03-0.05 at |z|max ~ 37) is above the literal 1e-2 bound, which is
        # why it is opt-in and 3xTF32 is the default
        assert e["theta"] < 1e-3 and e["grad"] < 1e-3 and e["logits"] < 0.1


def test_adaptation_from_pretrained_state_all_modes():
    """A non-degenerate state (oracle pre-trained on one task until it segments), then the canonical protocol on a
    second task: 5 Adam steps of batch 8 on 5 support images, transductive prediction of 5 query images.  Every
    numeric mode must agree with the float64 oracle on adapted weights, query logits and per-task mIoU."""
    from mliis_b200 import native as N
    from mliis_b200.synthetic import make_task_arrays, parse_records
    from oracle.efficientlab_oracle import BN_MOMENTUM
    size, B, T = 64, 8, 5
    arch, theta, bn, _, _ = make_problem(size, 2, task_id=0)
    pools = [parse_records(*make_task_arrays(t, 6, size)) for t in range(8)]          # pre-train on 8 tasks
    x0 = torch.from_numpy(np.concatenate([p[0] for p in pools]))
    y0 = torch.from_numpy(np.concatenate([p[1] for p in pools]))
    o32 = EfficientLabOracle(arch, torch.float32)
    th, bns = theta.float(), bn.float()
    opt = OptState(arch.n_params, torch.float32)
    g = torch.Generator().manual_seed(0)
    for s in range(100):
        idx = torch.randint(0, x0.shape[0], (8,), generator=g)
        _, gr, bns, _ = o32.loss_and_grad(th, bns, x0[idx], y0[idx])
        th = opt.apply(th, gr, 1e-3)
    # BN recalibration: the moving statistics (momentum 0.99) lag far behind after 100 steps; replace them by the
    # batch statistics of 16 images drawn across the pre-training tasks so that eval-mode predictions are meaningful
    o64 = EfficientLabOracle(arch, torch.float64)
    b0 = bns.double()
    _, nb = o64.forward(th.double(), b0, x0[torch.arange(0, 48, 3)], True)
    theta = th.double()
    bn = (b0 + (nb - b0) / (1 - BN_MOMENTUM)).float().double()
    images, labels = parse_records(*make_task_arrays(20, 10, size))                    # the held-out task
    orc = EfficientLabOracle(arch, torch.float64)
    rng = np.random.default_rng(0)
    batches = [rng.integers(0, 5, B).astype(np.int32) for _ in range(T)]
    tho, bno = theta, bn
    # the pre-trained "checkpoint" carries its optimizer slots (Gecko._full_state covers all global variables)
    oo = OptState(arch.n_params, torch.float64)
    oo.v = opt.v.double().clone()
    oo.b1p, oo.b2p = opt.b1p, float(np.float32(opt.b2p))
    v_slots, b2p0 = opt.v.double().clone(), oo.b2p
    for b in batches:
        _, gr, bno, _ = orc.loss_and_grad(tho, bno, torch.from_numpy(images[b]), torch.from_numpy(labels[b]))
        tho = oo.apply(tho, gr, 1e-3)
    pred_ref, lg_ref = orc.predict(tho, bno, torch.from_numpy(images[5:10]))
    iou_ref = np.mean([(lambda c: (c[0] + 1e-7) / (c[1] + 1e-7))(iou_counts(pred_ref.numpy()[j], labels[5 + j])) for j in range(5)])
    assert iou_ref > 0.2, "pre-training did not produce a segmenting model (iou %.3f)" % iou_ref
    xd, yd = _dev(images), _dev(labels)
    q = torch.arange(5, 10, dtype=torch.int32).cuda()
    rows = []
    for mode in (N.GEMM_FP32, N.GEMM_TF32X3, N.GEMM_TF32):
        eng = make_engine(arch, theta, bn, size, B, gemm_mode=mode)
        eng.init_state(0, split_vars(arch, theta), bn[0].numpy(), bn[1].numpy(), adam_v=split_vars(arch, v_slots),
                       beta1_power=0.0, beta2_power=b2p0)
        for b in batches:
            eng.train_step(0, xd, yd, 1e-3, index=torch.from_numpy(b).cuda())
        _, lg, inter, uni = eng.predict(0, xd, yd, index=q, want_pred=False, want_logits=True)
        torch.cuda.synchronize()
        e_theta = rel_l2(eng.tf_order_vector(eng.theta(0)).cpu().double(), tho)
        e_logits = (lg.cpu().double() - lg_ref).abs().max().item()
        iou = float(np.mean((inter.cpu().numpy() + 1e-7) / (uni.cpu().numpy() + 1e-7)))
        rows.append((mode, e_theta, e_logits, iou))
    msg = "pretrained-state adaptation: oracle mIoU %.4f |z|max %.1f ; (mode, theta relL2, logits maxabs, mIoU) = %s" % (
        iou_ref, lg_ref.abs().max().item(), rows)
    print(msg)
    try:
        with open("gpurun_out/tf32_parity.log", "a") as f:
            f.write(msg + "\n")
    except OSError:
        pass
    # NB the post-adaptation logits differ from the float64 oracle by ~0.1 (|z|max ~ 37) in EVERY mode, fp32
    # included: five Adam(beta1=0) steps amplify rounding-level gradient differences; the
    # 1e-2 logits bound is asserted on same-weights forwards (test_forward_layers, test_predict_mask_and_iou_counts,
    # test_tensor_core_modes_*), here the north-star bounds on adapted weights and per-task mIoU are asserted.
    for mode, e_theta, e_logits, iou in rows:
        assert e_theta < 1e-3, (mode, e_theta)
        assert abs(iou - iou_ref) < 0.005, (mode, iou, iou_ref)
        assert e_logits < 0.5, (mode, e_logits)


def test_engine_against_committed_oracle_fixture():
    """The CUDA path against the COMMITTED fixture tests/golden/oracle_small.npz (float64 oracle outputs on a seeded
    32x32 problem; generated by tests/golden/make_golden.py), without running the oracle."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz"))
    arch, theta, bn, images, labels = make_problem(32, 2, task_id=3, theta_seed=1)
    for mode, tol in ((0, 1e-4), (2, 3e-4)):
        eng = make_engine(arch, theta, bn, 32, 2, gemm_mode=mode)
        xd, yd = _dev(images), _dev(labels)
        logits = eng.forward(0, xd, True)
        loss, grads = eng.loss_backward(0, yd, 2)
        torch.cuda.synchronize()
        assert abs(loss.item() - float(gold["loss"])) < 1e-4 * max(1.0, abs(float(gold["loss"])))
        assert np.abs(logits.cpu().numpy()[:, ::4, ::4, :] - gold["logits"]).max() < 1e-3
        sel = np.arange(0, arch.n_params, 997)
        g = eng.tf_order_vector(grads).cpu().double().numpy()
        assert np.linalg.norm(g[sel] - gold["grad_sel"]) < tol * 10 * np.linalg.norm(gold["grad_sel"])
        assert abs(np.linalg.norm(g) - float(gold["grad_norm"])) < 1e-3 * float(gold["grad_norm"])
        eng.optimizer_step(0, 1e-3)
        torch.cuda.synchronize()
        th1 = eng.tf_order_vector(eng.theta(0)).cpu().double().numpy()
        assert np.linalg.norm(th1[sel] - gold["theta1_sel"]) < 1e-3 * np.linalg.norm(gold["theta1_sel"])
        assert np.abs(eng.bn_state(0)[0, :64].cpu().numpy() - gold["bn_mean_head"]).max() < 1e-5
