"""Per-kernel C-ABI entry points of the HBM-bound kernels and of the folded decoder conv against
float64 torch references, alone and task-batched (mliis_kernel_group: one launch serves several slot copies)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.efficientlab_oracle import (BN_EPS, BN_MOMENTUM, conv2d_same, depthwise_same, resize_bilinear_ac, swish)
from tests.parity_util import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _lib():
    from mliis_b200 import native as N
    return N, N.lib()


class Arena:
    """n slot copies of a set of named buffers at one uniform stride (what task-batched launches address)."""

    def __init__(self, n, spec):
        self.n, self.off, o = n, {}, 0
        for name, numel in spec.items():
            self.off[name] = (o, numel)
            o += (numel + 63) // 64 * 64
        self.stride = o
        self.buf = torch.zeros(n, o, dtype=torch.float32, device="cuda")

    def t(self, name, slot=0):
        o, numel = self.off[name]
        return self.buf[slot, o:o + numel]

    def p(self, name):
        return self.t(name).data_ptr()

    @property
    def stride_bytes(self):
        return self.stride * 4


def _bn_ref(x, gamma, beta):
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    rstd = torch.rsqrt(var + BN_EPS)
    a = gamma * rstd
    return mean, var, rstd, a, beta - mean * a


@pytest.mark.parametrize("k,stride,H,C", [(3, 1, 28, 32), (5, 2, 28, 48), (3, 2, 14, 96), (5, 1, 14, 40)])
def test_dwconv_bwd(k, stride, H, C):
    N, lib = _lib()
    g = torch.Generator().manual_seed(k * 10 + stride + H)
    B, n = 2, 3
    Ho = (H + stride - 1) // stride
    scratch = int(lib.mliis_kernel_scratch_floats(B, H, H, C))
    ar = Arena(n, dict(x=B * H * H * C, a=C, b=C, w=k * k * C, dy=B * Ho * Ho * C, dx=B * H * H * C, dw=k * k * C, s=scratch))
    refs = []
    for s in range(n):
        x = torch.randn(B, H, H, C, generator=g, dtype=torch.float64, requires_grad=True)
        w = (torch.randn(k, k, C, 1, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
        a = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
        b = torch.randn(C, generator=g, dtype=torch.float64) * 0.2
        dy = torch.randn(B, Ho, Ho, C, generator=g, dtype=torch.float64)
        act = swish(x * a + b)
        act.retain_grad()
        y = depthwise_same(act.permute(0, 3, 1, 2), w, stride).permute(0, 2, 3, 1)
        y.backward(dy)
        refs.append((act.grad, w.grad.reshape(k * k, C)))
        for name, v in (("x", x), ("a", a), ("b", b), ("w", w), ("dy", dy)):
            ar.t(name, s).copy_(v.detach().reshape(-1).float())
    for group in (1, n):
        ar.buf[:, ar.off["dx"][0]:].zero_()
        N.check(lib.mliis_kernel_group(group, ar.stride_bytes))
        for s in (range(n) if group == 1 else [0]):
            base = s * ar.stride_bytes
            N.check(lib.mliis_dwconv_bwd(ar.p("x") + base, ar.p("a") + base, ar.p("b") + base, ar.p("w") + base,
                                         ar.p("dy") + base, ar.p("dx") + base, ar.p("dw") + base, ar.p("s") + base, B, H, H,
                                         C, k, stride, None))
        N.check(lib.mliis_kernel_group(1, 0))
        torch.cuda.synchronize()
        for s in range(n):
            assert rel_err(ar.t("dx", s).view(B, H, H, C), refs[s][0]) < 2e-5, (group, s)
            assert rel_err(ar.t("dw", s).view(k * k, C), refs[s][1]) < 5e-5, (group, s)


@pytest.mark.parametrize("M,C,fused", [(1568, 240, 0), (6272, 112, 1), (300, 16, 0)])
def test_bn_stats_fwd_and_swish_bwd(M, C, fused):
    N, lib = _lib()
    g = torch.Generator().manual_seed(M + C)
    n = 2
    scratch = int(lib.mliis_kernel_scratch_floats(1, 1, M, C))
    ar = Arena(n, dict(x=M * C, gamma=C, beta=C, mm=C, mv=C, stats=4 * C, g=M * C, dx=M * C, dgamma=C, dbeta=C, s=scratch))
    refs = []
    for s in range(n):
        x = (torch.randn(M, C, generator=g, dtype=torch.float64) * 2 + 0.5).requires_grad_(True)
        gamma = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
        beta = (torch.randn(C, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
        mm0 = torch.randn(C, generator=g, dtype=torch.float64)
        mv0 = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
        gr = torch.randn(M, C, generator=g, dtype=torch.float64)
        mean, var, rstd, a, b = _bn_ref(x, gamma, beta)
        y = swish(x * a + b)
        y.backward(gr)
        var_ema = var * (M / (M - 1.0)) if fused else var
        refs.append(dict(mean=mean.detach(), rstd=rstd.detach(), a=a.detach(), b=b.detach(),
                         mm=mm0 - (mm0 - mean.detach()) * (1 - BN_MOMENTUM), mv=mv0 - (mv0 - var_ema.detach()) * (1 - BN_MOMENTUM),
                         dx=x.grad, dgamma=gamma.grad, dbeta=beta.grad))
        for name, v in (("x", x), ("gamma", gamma), ("beta", beta), ("mm", mm0), ("mv", mv0), ("g", gr)):
            ar.t(name, s).copy_(v.detach().reshape(-1).float())
    N.check(lib.mliis_kernel_group(n, ar.stride_bytes))
    N.check(lib.mliis_bn_stats_fwd(ar.p("x"), ar.p("gamma"), ar.p("beta"), ar.p("mm"), ar.p("mv"), ar.p("stats"), ar.p("s"),
                                   M, C, fused, None))
    N.check(lib.mliis_bn_swish_bwd(ar.p("x"), ar.p("g"), ar.p("dx"), ar.p("stats"), ar.p("gamma"), ar.p("dgamma"),
                                   ar.p("dbeta"), ar.p("s"), M, C, None))
    N.check(lib.mliis_kernel_group(1, 0))
    torch.cuda.synchronize()
    for s in range(n):
        st = ar.t("stats", s).view(4, C)
        r = refs[s]
        assert rel_err(st[0], r["mean"]) < 1e-5 and rel_err(st[1], r["rstd"]) < 1e-5
        assert rel_err(st[2], r["a"]) < 1e-5 and rel_err(st[3], r["b"]) < 1e-5
        assert rel_err(ar.t("mm", s), r["mm"]) < 1e-6 and rel_err(ar.t("mv", s), r["mv"]) < 1e-6
        assert rel_l2(ar.t("dx", s).view(M, C), r["dx"]) < 2e-5
        assert rel_err(ar.t("dgamma", s), r["dgamma"]) < 5e-5 and rel_err(ar.t("dbeta", s), r["dbeta"]) < 5e-5


def test_se_fwd():
    N, lib = _lib()
    g = torch.Generator().manual_seed(9)
    B, HW, C, Cr, n = 4, 196, 480, 20, 2
    scratch = int(lib.mliis_kernel_scratch_floats(B, 14, 14, C))
    ar = Arena(n, dict(x=B * HW * C, a=C, b=C, w1=C * Cr, b1=Cr, w2=Cr * C, b2=C, pool=B * C, hid=B * Cr, gate=B * C, s=scratch))
    refs = []
    for s in range(n):
        x = torch.randn(B, HW, C, generator=g, dtype=torch.float64)
        a = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
        b = torch.randn(C, generator=g, dtype=torch.float64) * 0.2
        w1 = torch.randn(C, Cr, generator=g, dtype=torch.float64) * 0.1
        b1 = torch.randn(Cr, generator=g, dtype=torch.float64) * 0.1
        w2 = torch.randn(Cr, C, generator=g, dtype=torch.float64) * 0.3
        b2 = torch.randn(C, generator=g, dtype=torch.float64) * 0.1
        pool = swish(x * a + b).mean(1)
        gate = torch.sigmoid(swish(pool @ w1 + b1) @ w2 + b2)
        refs.append((pool, gate))
        for name, v in (("x", x), ("a", a), ("b", b), ("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2)):
            ar.t(name, s).copy_(v.reshape(-1).float())
    N.check(lib.mliis_kernel_group(n, ar.stride_bytes))
    N.check(lib.mliis_se_fwd(ar.p("x"), ar.p("a"), ar.p("b"), ar.p("w1"), ar.p("b1"), ar.p("w2"), ar.p("b2"), ar.p("pool"),
                             ar.p("hid"), ar.p("gate"), ar.p("s"), B, HW, C, Cr, None))
    N.check(lib.mliis_kernel_group(1, 0))
    torch.cuda.synchronize()
    for s in range(n):
        assert rel_err(ar.t("pool", s).view(B, C), refs[s][0]) < 1e-5
        assert rel_err(ar.t("gate", s).view(B, C), refs[s][1]) < 1e-5


@pytest.mark.parametrize("dice,ls", [(1, 0.0), (0, 0.1)])
def test_softmax_ce_iou(dice, ls):
    N, lib = _lib()
    g = torch.Generator().manual_seed(4 + dice)
    B, h, H, n = 3, 16, 64, 2
    scratch = int(lib.mliis_kernel_scratch_floats(B, H, H, 4))
    ar = Arena(n, dict(z=B * h * h * 2, y=B * H * H * 2, p1=B * H * H, dz=B * H * H * 2, s=scratch, loss=4))
    refs = []
    for s in range(n):
        z = (torch.randn(B, h, h, 2, generator=g, dtype=torch.float64) * 2).requires_grad_(True)
        fg = (torch.rand(B, H, H, generator=g) > 0.6).double()
        y = torch.stack([1 - fg, fg], -1)
        zh = resize_bilinear_ac(z.permute(0, 3, 1, 2), H, H).permute(0, 2, 3, 1)
        zh.retain_grad()
        logp = F.log_softmax(zh, dim=-1)
        yc = y * (1 - ls) + ls / 2 if ls > 0 else y
        loss = -(yc * logp).sum(-1).mean()
        if dice:
            p1 = torch.softmax(zh, -1)[..., 1].reshape(B, -1)
            y1 = y[..., 1].reshape(B, -1)
            inter = (p1 * y1).sum(1)
            iou = ((inter + 1e-7) / (p1.sum(1) + y1.sum(1) - inter + 1e-7)).mean()
            loss = loss - torch.log(2 * iou / (iou + 1))
        loss.backward()
        refs.append((loss.detach(), zh.grad))
        ar.t("z", s).copy_(z.detach().reshape(-1).float())
        ar.t("y", s).copy_(y.reshape(-1).float())
    N.check(lib.mliis_kernel_group(n, ar.stride_bytes))
    N.check(lib.mliis_softmax_ce_iou(ar.p("z"), ar.p("y"), ar.p("p1"), ar.p("dz"), ar.p("s"), ar.p("loss"), B, h, h, H, H, dice,
                                     ls, None))
    N.check(lib.mliis_kernel_group(1, 0))
    torch.cuda.synchronize()
    for s in range(n):
        assert abs(ar.t("loss", s)[0].item() - refs[s][0].item()) < 1e-5 * max(1.0, abs(refs[s][0].item()))
        assert rel_l2(ar.t("dz", s).view(B, H, H, 2), refs[s][1]) < 2e-5


@pytest.mark.parametrize("H,Cin,Cp,B", [(56, 224, 136, 2), (14, 224, 224, 3), (8, 224, 40, 2)])
def test_rsd_conv2_folded_forward(H, Cin, Cp, B):
    """conv2d_2 with the pooled branch folded == the reference's conv over concat([branches, tiled image mean])."""
    N, lib = _lib()
    g = torch.Generator().manual_seed(H + Cp)
    Cout, n = 112, 2
    ar = Arena(n, dict(x=B * H * H * Cin, pooled=B * Cp, w=9 * (Cin + Cp) * Cout, wt=2 * 9 * Cin * Cout, bias=Cout,
                       b9=B * 9 * Cout, y=B * H * H * Cout))
    refs = []
    for s in range(n):
        x = torch.randn(B, H, H, Cin, generator=g, dtype=torch.float64)
        pooled = torch.randn(B, Cp, generator=g, dtype=torch.float64)
        w = torch.randn(3, 3, Cin + Cp, Cout, generator=g, dtype=torch.float64) * 0.05
        bias = torch.randn(Cout, generator=g, dtype=torch.float64)
        full = torch.cat([x, pooled.view(B, 1, 1, Cp).expand(B, H, H, Cp)], -1)
        refs.append(conv2d_same(full.permute(0, 3, 1, 2), w, bias=bias).permute(0, 2, 3, 1))
        for name, v in (("x", x), ("pooled", pooled), ("w", w), ("bias", bias)):
            ar.t(name, s).copy_(v.reshape(-1).float())
    N.check(lib.mliis_kernel_group(n, ar.stride_bytes))
    N.check(lib.mliis_tc_prep_weights_sub(ar.p("w"), ar.p("wt"), 9, Cin, Cin + Cp, Cout, 0, N.GEMM_TF32X3, None))
    N.check(lib.mliis_rsd_conv2_fwd(ar.p("x"), Cin, ar.p("pooled"), ar.p("w"), ar.p("wt"), ar.p("bias"), ar.p("b9"), ar.p("y"),
                                    B, H, H, Cin, Cp, Cout, N.GEMM_TF32X3, None))
    N.check(lib.mliis_kernel_group(1, 0))
    torch.cuda.synchronize()
    for s in range(n):
        assert rel_err(ar.t("y", s).view(B, H, H, Cout), refs[s]) < 1e-4, s


@pytest.mark.parametrize("HW,Cin,Cout,B", [(56 * 56, 144, 24, 2), (14 * 14, 672, 112, 3), (49, 96, 16, 4),
                                           (112 * 112, 96, 24, 3), (100 * 100 + 3, 40, 136, 4)])
def test_project_conv_with_fused_prologue(HW, Cin, Cout, B):
    """mliis_tc_project_conv == (swish(a*x+b) * gate[img]) @ W   (efficientnet_model.py:225-232, :266, :271-273); M tiles
    that straddle two images (HW = 49) read both gates.  The two large-M cases run on the persistent tile-loop kernel
    (tc_pw_kernel: >= 2 row tiles per CTA), the last one with a ragged M, a K tail (40 = 32 + 8) and a non-wide N."""
    N, lib = _lib()
    g = torch.Generator().manual_seed(HW + Cin)
    n = 2
    ar = Arena(n, dict(x=B * HW * Cin, w=Cin * Cout, wt=2 * Cin * Cout, a=Cin, b=Cin, gate=B * Cin, y=B * HW * Cout))
    refs = []
    for s in range(n):
        x = torch.randn(B, HW, Cin, generator=g, dtype=torch.float64)
        w = torch.randn(Cin, Cout, generator=g, dtype=torch.float64) * 0.1
        a = torch.rand(Cin, generator=g, dtype=torch.float64) + 0.5
        b = torch.randn(Cin, generator=g, dtype=torch.float64) * 0.3
        gate = torch.rand(B, Cin, generator=g, dtype=torch.float64)
        refs.append((swish(x * a + b) * gate[:, None, :]) @ w)
        for name, v in (("x", x), ("w", w), ("a", a), ("b", b), ("gate", gate)):
            ar.t(name, s).copy_(v.reshape(-1).float())
    N.check(lib.mliis_kernel_group(n, ar.stride_bytes))
    N.check(lib.mliis_tc_prep_weights(ar.p("w"), ar.p("wt"), 1, Cin, Cout, 0, N.GEMM_TF32X3, None))
    N.check(lib.mliis_tc_project_conv(ar.p("x"), ar.p("wt"), ar.p("a"), ar.p("b"), ar.p("gate"), ar.p("y"), B, HW, Cin, Cout,
                                      N.GEMM_TF32X3, None))
    N.check(lib.mliis_kernel_group(1, 0))
    torch.cuda.synchronize()
    for s in range(n):
        assert rel_err(ar.t("y", s).view(B, HW, Cout), refs[s]) < 1e-4, s


@pytest.mark.parametrize("M,K,N,bias", [(37632, 16, 96, False), (40003, 24, 144, True), (6272, 24, 144, True)])
def test_pointwise_conv(M, K, N, bias):
    """mliis_tc_conv with taps = 1 (MBConv expand conv / any plain GEMM) on both pointwise kernels: the persistent tile
    loop (large M) and one CTA per tile (small M)."""
    Nn, lib = _lib()
    g = torch.Generator().manual_seed(M + K)
    n = 2
    ar = Arena(n, dict(x=M * K, w=K * N, wt=2 * K * N, b=N, y=M * N))
    refs = []
    for s in range(n):
        x = torch.randn(M, K, generator=g, dtype=torch.float64)
        w = torch.randn(K, N, generator=g, dtype=torch.float64) * 0.2
        b = torch.randn(N, generator=g, dtype=torch.float64)
        refs.append(x @ w + (b if bias else 0.0))
        for name, v in (("x", x), ("w", w), ("b", b)):
            ar.t(name, s).copy_(v.reshape(-1).float())
    Nn.check(lib.mliis_kernel_group(n, ar.stride_bytes))
    Nn.check(lib.mliis_tc_prep_weights(ar.p("w"), ar.p("wt"), 1, K, N, 0, Nn.GEMM_TF32X3, None))
    Nn.check(lib.mliis_tc_conv(ar.p("x"), ar.p("wt"), ar.p("b") if bias else None, ar.p("y"), 1, 1, M, K, N, 1, 1,
                               Nn.GEMM_TF32X3, None))
    Nn.check(lib.mliis_kernel_group(1, 0))
    torch.cuda.synchronize()
    for s in range(n):
        assert rel_err(ar.t("y", s).view(M, N), refs[s]) < 1e-4, s
