"""The oracle against independent checks (CPU).  PARITY UNPINNED: the reference has no golden vectors for the
float path, so the oracle is pinned by (a) finite differences, (b) independent torch primitives,
(c) hand-written formulas, (d) the alternative IoU metric kept in the reference, (e) its own committed regression
fixture tests/golden/oracle_small.npz."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.efficientlab_oracle import (ADAM_BETA2, ADAM_EPS, Arch, EfficientLabOracle, OptState, conv2d_same,
                                        depthwise_same, iou_counts, iou_score, resize_bilinear_ac, resize_tables,
                                        same_pad)
from tests.parity_util import make_problem

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_variable_tables_match_survey():
    a = Arch()
    assert a.n_params == 2071714 and len(a.params) == 169 and len(a.bns) == 39 and a.n_bn == 8752
    assert a.dc_blocks == [2, 4, 6, 7, 9, 10]
    assert a.reduction_block == {1: 0, 2: 2, 3: 4, 4: 10}
    assert sum(p.size for p in a.params if p.name.startswith("efficientnet-b0")) == 851808
    assert sum(p.size for p in a.params if p.name.startswith("decode")) == 1219906
    assert abs(a.blocks[2].dc_rate - 0.2 * 2 / 11) < 1e-12 and abs(a.blocks[10].dc_rate - 0.2 * 10 / 11) < 1e-12
    names = {p.name for p in a.params}
    for n in ["efficientnet-b0/model/stem/conv2d/kernel", "efficientnet-b0/model/blocks_0/conv2d/kernel",
              "efficientnet-b0/model/blocks_1/conv2d_1/kernel", "efficientnet-b0/model/blocks_10/se/conv2d_1/bias",
              "efficientnet-b0/model/blocks_3/tpu_batch_normalization_2/gamma",
              "decode/decode_skip_connections_3/conv2d_2/kernel", "decode/decode_skip_connections_1/batch_normalization_1/beta",
              "decode/final_layer_weights/bias"]:
        assert n in names, n
    # l2_term excludes every variable whose name contains 'batch_normalization' (regularizers.py:9)
    assert all(p.l2 == ("batch_normalization" not in p.name) for p in a.params)


def test_same_padding_rule():
    # asymmetric for stride 2: stem / blocks_1,5 pad (0,1); blocks_3 (5x5 s2 on 56) pads (1,2)
    assert same_pad(224, 3, 2) == (0, 1)
    assert same_pad(112, 3, 2) == (0, 1)
    assert same_pad(56, 5, 2) == (1, 2)
    assert same_pad(28, 3, 2) == (0, 1)
    assert same_pad(14, 5, 1) == (2, 2)
    assert same_pad(56, 3, 1, 2) == (2, 2)       # dilation 2
    # output size is ceil(n/s)
    x = torch.randn(1, 3, 11, 11, dtype=torch.float64)
    w = torch.randn(3, 3, 3, 4, dtype=torch.float64)
    assert conv2d_same(x, w, stride=2).shape[-1] == 6
    wd = torch.randn(5, 5, 3, 1, dtype=torch.float64)
    assert depthwise_same(x, wd, 2).shape[-1] == 6


def test_conv_matches_explicit_loops():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 2, 5, 5, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 2, 3, generator=g, dtype=torch.float64)
    y = conv2d_same(x, w, stride=2)
    lo, _ = same_pad(5, 3, 2)
    ref = torch.zeros(1, 3, 3, 3, dtype=torch.float64)
    for oy in range(3):
        for ox in range(3):
            for ky in range(3):
                for kx in range(3):
                    iy, ix = oy * 2 - lo + ky, ox * 2 - lo + kx
                    if 0 <= iy < 5 and 0 <= ix < 5:
                        ref[0, :, oy, ox] += x[0, :, iy, ix] @ w[ky, kx]
    assert torch.allclose(y, ref, atol=1e-12)


def test_bilinear_align_corners_cross_check():
    g = torch.Generator().manual_seed(1)
    for (hi, ho) in [(14, 56), (56, 224), (4, 16), (7, 7)]:
        x = torch.randn(2, 3, hi, hi, generator=g, dtype=torch.float64)
        ours = resize_bilinear_ac(x, ho, ho)
        ref = F.interpolate(x, size=(ho, ho), mode="bilinear", align_corners=True)
        assert (ours - ref).abs().max() < 5e-5        # the tables are float32 like TF's
    lo, hi_, lerp = resize_tables(14, 56)
    assert lo[0] == 0 and hi_[-1] == 13 and lerp[0] == 0 and abs(lerp[-1]) < 1e-6 and lo[-1] == 13


def test_adam_matches_tf_formula_and_differs_from_torch_eps_placement():
    n = 50
    g = torch.Generator().manual_seed(2)
    th = torch.randn(n, generator=g, dtype=torch.float64)
    gr = torch.randn(n, generator=g, dtype=torch.float64) * 1e-3
    st = OptState(n, torch.float64)
    out = st.apply(th, gr, 1e-3)
    # first step, beta1 = 0:  v = (1-b2) g^2 ; alpha = lr*sqrt(1-b2) ; step = alpha*g/(sqrt(v)+eps) ~ lr*sign(g)
    v = (1 - ADAM_BETA2) * gr * gr
    ref = th - 1e-3 * math.sqrt(1 - ADAM_BETA2) * gr / (torch.sqrt(v) + ADAM_EPS)
    assert torch.allclose(out, ref, atol=1e-15)
    assert abs(st.b2p - ADAM_BETA2 ** 2) < 1e-15 and st.b1p == 0.0
    big = gr.abs() > 5e-4
    assert torch.allclose((th - out)[big], 1e-3 * torch.sign(gr)[big], rtol=2e-3)
    # in the eps -> 0 limit it coincides with torch.optim.Adam
    p = th.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.0, 0.999), eps=1e-30)
    p.grad = gr.clone()
    opt.step()
    st2 = OptState(n, torch.float64)
    import oracle.efficientlab_oracle as O
    old = O.ADAM_EPS
    O.ADAM_EPS = 1e-30
    try:
        out2 = st2.apply(th, gr, 1e-3)
    finally:
        O.ADAM_EPS = old
    assert torch.allclose(out2, p.detach(), atol=1e-12)


def test_bn_train_and_ema_rules():
    arch = Arch()
    orc = EfficientLabOracle(arch, torch.float64)
    theta = arch.init_theta(0)
    bn = arch.init_bn_state()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 32, 6, 6, generator=g, dtype=torch.float64) * 2 + 1
    new_bn = bn.clone()
    y = orc._bn(x, theta, bn, new_bn, "efficientnet-b0/model/stem/tpu_batch_normalization", True)
    ref = F.batch_norm(x, None, None, None, None, True, 0.0, 1e-3)
    assert torch.allclose(y, ref, atol=1e-10)
    mean, var = x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False)
    assert torch.allclose(new_bn[0, :32], 0.01 * mean, atol=1e-12)                 # backbone: biased variance
    assert torch.allclose(new_bn[1, :32], 1 - 0.01 * (1 - var), atol=1e-12)
    spec = arch.bn_by_name["decode/decode_skip_connections_3/batch_normalization"]
    x2 = torch.randn(4, 112, 3, 3, generator=g, dtype=torch.float64)
    nb2 = bn.clone()
    orc._bn(x2, theta, bn, nb2, spec.name, True)
    n = 4 * 9
    var_u = x2.var((0, 2, 3), unbiased=False) * n / (n - 1)                          # decoder: Bessel corrected
    assert torch.allclose(nb2[1, spec.offset:spec.offset + 112], 1 - 0.01 * (1 - var_u), atol=1e-12)


def test_gradient_finite_differences():
    arch, theta, bn, images, labels = make_problem(32, 2)
    orc = EfficientLabOracle(arch, torch.float64)
    x, y = torch.from_numpy(images), torch.from_numpy(labels)
    dc = torch.tensor([[1, 0], [1, 1], [0, 1], [1, 1], [1, 0], [0, 0]], dtype=torch.float64)
    loss, g, _, _ = orc.loss_and_grad(theta, bn, x, y, dc_masks=dc)
    rng = np.random.default_rng(0)
    # probe one coordinate of a spread of tensors (kernels, depthwise, SE, BN, decoder, head)
    names = ["efficientnet-b0/model/stem/conv2d/kernel", "efficientnet-b0/model/blocks_1/depthwise_conv2d/depthwise_kernel",
             "efficientnet-b0/model/blocks_4/se/conv2d/kernel", "efficientnet-b0/model/blocks_6/tpu_batch_normalization_1/gamma",
             "efficientnet-b0/model/blocks_9/conv2d_1/kernel", "decode/decode_skip_connections_3/conv2d_2/kernel",
             "decode/decode_skip_connections_1/conv2d_1/bias", "decode/decode_skip_connections_1/batch_normalization/beta",
             "decode/final_layer_weights/kernel"]
    for nme in names:
        p = arch.by_name[nme]
        i = p.offset + int(rng.integers(p.size))
        eps = 1e-5

        def f(delta):
            th = theta.clone()
            th[i] += delta
            logits, _ = orc.forward(th, bn, x, True, dc_masks=dc)
            return orc.loss(th, logits, y).item()
        fd = (f(eps) - f(-eps)) / (2 * eps)
        assert abs(fd - g[i].item()) < 1e-6 + 1e-4 * abs(fd), (nme, fd, g[i].item())


def test_iou_metric_cross_check():
    rng = np.random.default_rng(4)
    for _ in range(5):
        pred = (rng.random((16, 16, 2)) > 0.5).astype(np.float32)
        lab1 = (rng.random((16, 16)) > 0.4).astype(np.float32)
        lab = np.stack([1 - lab1, lab1], -1)
        i, u = iou_counts(pred, lab)
        tp = np.logical_and(lab1 > 0.5, pred[..., 1] > 0.5).sum()          # reptile.py:555-566 (measure / iou_img)
        fp = np.logical_and(lab1 <= 0.5, pred[..., 1] > 0.5).sum()
        fn = np.logical_and(lab1 > 0.5, pred[..., 1] <= 0.5).sum()
        assert i == tp and u == tp + fp + fn
        assert abs(iou_score(pred, lab) - tp / max(tp + fp + fn, 1)) < 1e-6
    z = np.zeros((4, 4, 2), np.float32)
    assert iou_score(z, z) == 1.0                      # (0+eps)/(0+eps), reptile.py:549


def test_oracle_regression_fixture():
    from oracle.efficientlab_oracle import OptState
    gold = np.load(os.path.join(GOLD, "oracle_small.npz"))
    arch, theta, bn, images, labels = make_problem(32, 2, task_id=3, theta_seed=1)
    orc = EfficientLabOracle(arch, torch.float64)
    loss, g, nbn, logits = orc.loss_and_grad(theta, bn, torch.from_numpy(images), torch.from_numpy(labels))
    sel = np.arange(0, arch.n_params, 997)
    assert abs(loss.item() - float(gold["loss"])) < 1e-9
    assert np.allclose(logits.numpy()[:, ::4, ::4, :], gold["logits"], atol=1e-9)
    assert np.allclose(g.numpy()[sel], gold["grad_sel"], atol=1e-10)
    th1 = OptState(arch.n_params, torch.float64).apply(theta, g, 1e-3)
    assert np.allclose(th1.numpy()[sel], gold["theta1_sel"], atol=1e-9)
    assert np.allclose(nbn[0, :64].numpy(), gold["bn_mean_head"], atol=1e-12)


def test_float32_mode_close_to_float64():
    arch, theta, bn, images, labels = make_problem(32, 2)
    o64 = EfficientLabOracle(arch, torch.float64)
    o32 = EfficientLabOracle(arch, torch.float32)
    x, y = torch.from_numpy(images), torch.from_numpy(labels)
    l64, g64, _, _ = o64.loss_and_grad(theta, bn, x, y)
    l32, g32, _, _ = o32.loss_and_grad(theta.float(), bn.float(), x, y)
    assert abs(l64.item() - l32.item()) < 1e-4
    assert ((g32.double() - g64).norm() / g64.norm()).item() < 1e-3


def test_multiclass_loss_restatement():
    """The joint-training loss (efficientlab.py:294-327 with binary_iou_loss=False, :369-396): hand computation from
    the formulas, the reduction to sparse labels used by the CUDA kernels, and a finite-difference check."""
    torch.manual_seed(0)
    B, H, C = 2, 6, 5
    arch = Arch(n_out=C)
    orc = EfficientLabOracle(arch, torch.float64, binary_iou_loss=False, l2=False, label_smoothing=0.1)
    logits = torch.randn(B, H, H, C, dtype=torch.float64)
    fg = (torch.rand(B, H, H) > 0.5)
    cls = torch.tensor([2, 4])
    t = torch.where(fg, cls[:, None, None].expand(B, H, H), torch.zeros(B, H, H, dtype=torch.long))
    y = F.one_hot(t, C).double()
    theta = torch.zeros(arch.n_params, dtype=torch.float64)
    loss = orc.loss(theta, logits, y).item()
    # by hand, pixel by pixel
    p = torch.softmax(logits, -1)
    ce = 0.0
    for b in range(B):
        for i in range(H):
            for j in range(H):
                ys = y[b, i, j] * 0.9 + 0.1 / C                      # tf.losses label_smoothing [TF-ext]
                ce -= float((ys * torch.log(p[b, i, j])).sum())
    ce /= B * H * H
    ious = []
    for b in range(B):
        inter = float((p[b] * y[b]).sum())
        den = float(p[b].sum() + y[b].sum()) - inter
        ious.append((inter + 1e-7) / (den + 1e-7))
    iou = sum(ious) / B
    assert abs(loss - (ce - np.log(2 * iou / (iou + 1)))) < 1e-12
    # sparse form used on the device: sum_c y = sum_c p = 1 per pixel  =>  IoU_b = (I_b + eps) / (2HW - I_b + eps)
    I = p.gather(-1, t[..., None])[..., 0].reshape(B, -1).sum(1)
    iou_sparse = ((I + 1e-7) / (2 * H * H - I + 1e-7)).mean().item()
    assert abs(iou_sparse - iou) < 1e-12
    # finite differences of d loss / d logits
    z = logits.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(orc.loss(theta, z, y), z)
    rng = np.random.default_rng(0)
    for _ in range(12):
        idx = tuple(int(rng.integers(0, n)) for n in logits.shape)
        e = torch.zeros_like(logits)
        e[idx] = 1e-6
        fd = (orc.loss(theta, logits + e, y) - orc.loss(theta, logits - e, y)).item() / 2e-6
        assert abs(fd - g[idx].item()) < 1e-7 + 1e-5 * abs(fd)
    # binary_iou_loss=True on a 2-channel problem scores channel 1 only: a different number
    arch2 = Arch()
    l_bin = EfficientLabOracle(arch2, torch.float64, l2=False).loss(torch.zeros(arch2.n_params, dtype=torch.float64),
                                                                   logits[..., :2], y[..., :2])
    l_all = EfficientLabOracle(arch2, torch.float64, l2=False, binary_iou_loss=False).loss(
        torch.zeros(arch2.n_params, dtype=torch.float64), logits[..., :2], y[..., :2])
    assert abs(l_bin.item() - l_all.item()) > 1e-3
