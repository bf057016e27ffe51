"""CPU restatement (torch-CPU, float64 master / float32 timing) of the EfficientLab graph
and of one ``session.run(minimize_op)`` / ``session.run(predictions)`` of the reference.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  PARITY UNPINNED (no reference golden
vectors exist; TF-1.15 cannot run here).

Reference files restated here (all paths into /root/reference):
  models/efficientlab.py:111-119   input normalisation
  models/efficientlab.py:126-231   decoder (residual skip decoder, head, resize, softmax)
  models/efficientlab.py:291-327   threshold, loss (CE - ln dice + L2), optimizer wiring
  models/efficientlab.py:329-396   soft IoU
  models/efficientnet/efficientnet_builder.py:125-149, :90-109   block strings, truncation
  models/efficientnet/efficientnet_model.py:133-290, :396-440    MBConv, stem, endpoints
  models/efficientnet/utils.py:87-134, :157-170                  BN (non fused), drop-connect
  models/regularizers.py:4-10                                    L2 term
  meta_learners/args.py:151-154                                  optimizer choice

TF-1.15 op semantics encoded from documentation/knowledge ([TF-ext] in SURVEY.md):
SAME padding (asymmetric for stride 2), fused-vs-non-fused BN moving-variance update,
ResizeBilinear(align_corners=True), softmax_cross_entropy reduction, Keras Dropout,
ApplyAdam / ApplyGradientDescent.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# models/efficientnet/constants.py:1-2
MEAN_RGB = [0.485 * 255, 0.456 * 255, 0.406 * 255]
STDDEV_RGB = [0.229 * 255, 0.224 * 255, 0.225 * 255]
# efficientnet_builder.py:137-138 ; tf.layers.batch_normalization defaults are the same values
BN_MOMENTUM = 0.99
BN_EPS = 1e-3
DROP_CONNECT_RATE = 0.2          # efficientnet_builder.py:128 (never overridden)
L2_COEF = 0.0005                 # regularizers.py:4
ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.0, 0.999, 1e-8   # efficientlab.py:16 + TF defaults

# efficientnet_builder.py:130-135
_BLOCK_STRINGS = [
    'r1_k3_s11_e1_i32_o16_se0.25', 'r2_k3_s22_e6_i16_o24_se0.25',
    'r2_k5_s22_e6_i24_o40_se0.25', 'r3_k3_s22_e6_i40_o80_se0.25',
    'r3_k5_s11_e6_i80_o112_se0.25', 'r4_k5_s22_e6_i112_o192_se0.25',
    'r1_k3_s11_e6_i192_o320_se0.25',
]


@dataclass
class Block:
    kernel: int
    stride: int
    cin: int
    cout: int
    expand: int
    se_reduced: int
    skip: bool
    dc_rate: float


def decode_blocks(max_block_num: int = 10) -> List[Block]:
    """efficientnet_builder.py:90-109 (truncate at string granularity) +
    efficientnet_model.py:326-349 (repeat expansion) + :426-428 (drop-connect rate)."""
    stages = []
    num_blocks = 0
    for s in _BLOCK_STRINGS:
        opts = {}
        for op in s.split('_'):
            m = re.split(r'(\d.*)', op)
            if len(m) >= 2:
                opts[m[0]] = m[1]
        r = int(opts['r'])
        num_blocks += r
        if num_blocks > max_block_num + 1:
            break
        stages.append((r, int(opts['k']), int(opts['s'][0]), int(opts['e']), int(opts['i']), int(opts['o']),
                       float(opts['se'])))
    blocks: List[Block] = []
    for (r, k, s, e, i, o, se) in stages:
        for rep in range(r):
            cin = i if rep == 0 else o
            stride = s if rep == 0 else 1
            blocks.append(Block(k, stride, cin, o, e, max(1, int(cin * se)), stride == 1 and cin == o, 0.0))
    n = len(blocks)
    for idx, b in enumerate(blocks):
        b.dc_rate = DROP_CONNECT_RATE * float(idx) / n
    return blocks


@dataclass
class ParamSpec:
    name: str            # expected TF variable name
    shape: Tuple[int, ...]
    offset: int          # into the oracle's flat theta (TF creation order)
    l2: bool             # regularizers.py:9 : name lacks 'batch_normalization'
    init: str            # 'conv' | 'glorot' | 'zeros' | 'ones'

    @property
    def size(self) -> int:
        return int(np.prod(self.shape))


@dataclass
class BNSpec:
    name: str            # scope of the BN layer
    channels: int
    offset: int          # channel offset into the flat moving_mean / moving_variance
    fused: bool          # decoder tf.layers.batch_normalization => FusedBatchNorm (Bessel-corrected EMA)


class Arch:
    """Variable tables of EfficientLab(feature_extractor='efficientnet-b0', rsd=[2,4])."""

    def __init__(self, rsd=(2, 4), aspp_dim: int = 112, n_out: int = 2, max_block_num: int = 10):
        self.blocks = decode_blocks(max_block_num)
        self.rsd = sorted(rsd, reverse=True)     # efficientlab.py:155
        self.D = aspp_dim
        self.n_out = n_out
        self.params: List[ParamSpec] = []
        self.bns: List[BNSpec] = []
        self._poff = 0
        self._boff = 0
        pre = 'efficientnet-b0/model/'
        self._conv(pre + 'stem/conv2d/kernel', (3, 3, 3, 32))
        self._bn(pre + 'stem/tpu_batch_normalization', 32, False)
        for i, b in enumerate(self.blocks):
            sc = pre + 'blocks_%d/' % i
            ce = b.cin * b.expand
            nconv = 0
            nbn = 0

            def cname(n):
                return 'conv2d' if n == 0 else 'conv2d_%d' % n

            def bname(n):
                return 'tpu_batch_normalization' if n == 0 else 'tpu_batch_normalization_%d' % n
            if b.expand != 1:
                self._conv(sc + cname(nconv) + '/kernel', (1, 1, b.cin, ce)); nconv += 1
                self._bn(sc + bname(nbn), ce, False); nbn += 1
            self._conv(sc + 'depthwise_conv2d/depthwise_kernel', (b.kernel, b.kernel, ce, 1))
            self._bn(sc + bname(nbn), ce, False); nbn += 1
            self._conv(sc + 'se/conv2d/kernel', (1, 1, ce, b.se_reduced))
            self._add(sc + 'se/conv2d/bias', (b.se_reduced,), True, 'zeros')
            self._conv(sc + 'se/conv2d_1/kernel', (1, 1, b.se_reduced, ce))
            self._add(sc + 'se/conv2d_1/bias', (ce,), True, 'zeros')
            self._conv(sc + cname(nconv) + '/kernel', (1, 1, ce, b.cout)); nconv += 1
            self._bn(sc + bname(nbn), b.cout, False); nbn += 1
        # reduction endpoints: efficientnet_model.py:417-439
        self.reduction_block = {}
        ridx = 0
        for i, b in enumerate(self.blocks):
            if i == len(self.blocks) - 1 or self.blocks[i + 1].stride > 1:
                ridx += 1
                self.reduction_block[ridx] = i
        # decoder: efficientlab.py:155-159, :179-231
        deep_c = self.blocks[self.reduction_block[4]].cout
        for r in self.rsd:
            skip_c = self.blocks[self.reduction_block[r]].cout
            sc = 'decode/decode_skip_connections_%d/' % (r - 1)
            assert deep_c == self.D, "extra 1x1 branch (efficientlab.py:213-215) not needed for b0"
            cat = deep_c + skip_c
            self._add(sc + 'conv2d/kernel', (1, 1, cat, self.D), True, 'glorot')
            self._add(sc + 'conv2d/bias', (self.D,), True, 'zeros')
            self._bn(sc + 'batch_normalization', self.D, True)
            self._add(sc + 'conv2d_1/kernel', (3, 3, cat, self.D), True, 'glorot')
            self._add(sc + 'conv2d_1/bias', (self.D,), True, 'zeros')
            self._bn(sc + 'batch_normalization_1', self.D, True)
            self._add(sc + 'conv2d_2/kernel', (3, 3, 2 * self.D + cat, self.D), True, 'glorot')
            self._add(sc + 'conv2d_2/bias', (self.D,), True, 'zeros')
            self._bn(sc + 'batch_normalization_2', self.D, True)
            deep_c = self.D
        self._conv('decode/final_layer_weights/kernel', (1, 1, self.D, self.n_out))
        self._add('decode/final_layer_weights/bias', (self.n_out,), True, 'zeros')
        self.n_params = self._poff
        self.n_bn = self._boff
        self.by_name = {p.name: p for p in self.params}
        self.bn_by_name = {b.name: b for b in self.bns}
        self.dc_blocks = [i for i, b in enumerate(self.blocks) if b.skip and b.dc_rate > 0]

    def _add(self, name, shape, l2, init):
        p = ParamSpec(name, tuple(shape), self._poff, l2, init)
        self.params.append(p)
        self._poff += p.size

    def _conv(self, name, shape):
        self._add(name, shape, True, 'conv')

    def _bn(self, scope, c, fused):
        self._add(scope + '/gamma', (c,), False, 'ones')
        self._add(scope + '/beta', (c,), False, 'zeros')
        self.bns.append(BNSpec(scope, c, self._boff, fused))
        self._boff += c

    # ---- initialisation (efficientnet_model.py:61-82; tf.layers default glorot_uniform) ----
    def init_theta(self, seed: int = 0, dtype=torch.float64) -> torch.Tensor:
        g = torch.Generator().manual_seed(seed)
        theta = torch.zeros(self.n_params, dtype=torch.float64)
        for p in self.params:
            v = theta[p.offset:p.offset + p.size]
            if p.init == 'conv':
                kh, kw, _, co = p.shape
                v.copy_(torch.randn(p.size, generator=g, dtype=torch.float64) * math.sqrt(2.0 / (kh * kw * co)))
            elif p.init == 'glorot':
                kh, kw, ci, co = p.shape
                lim = math.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
                v.copy_((torch.rand(p.size, generator=g, dtype=torch.float64) * 2 - 1) * lim)
            elif p.init == 'ones':
                v.fill_(1.0)
        return theta.to(dtype)

    def init_bn_state(self, dtype=torch.float64) -> torch.Tensor:
        """[2, n_bn]: row 0 moving_mean (zeros), row 1 moving_variance (ones)."""
        s = torch.zeros(2, self.n_bn, dtype=dtype)
        s[1].fill_(1.0)
        return s

    def l2_mask(self, dtype=torch.float64) -> torch.Tensor:
        m = torch.zeros(self.n_params, dtype=dtype)
        for p in self.params:
            if p.l2:
                m[p.offset:p.offset + p.size] = 1.0
        return m


# --------------------------------------------------------------------------------------
# TF op restatements
# --------------------------------------------------------------------------------------

def same_pad(n: int, k: int, s: int, d: int = 1) -> Tuple[int, int]:
    """TF 'SAME' padding [TF-ext]: total = max((ceil(n/s)-1)*s + (k-1)*d+1 - n, 0); lo = total//2."""
    out = -(-n // s)
    keff = (k - 1) * d + 1
    p = max((out - 1) * s + keff - n, 0)
    return p // 2, p - p // 2


def conv2d_same(x, w_hwio, stride=1, dilation=1, bias=None):
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb = same_pad(x.shape[2], kh, stride, dilation)
    pl, pr = same_pad(x.shape[3], kw, stride, dilation)
    x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1), bias, stride=stride, dilation=dilation)


def depthwise_same(x, w_hwc1, stride=1):
    kh, kw, c, _ = w_hwc1.shape
    pt, pb = same_pad(x.shape[2], kh, stride)
    pl, pr = same_pad(x.shape[3], kw, stride)
    x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w_hwc1.permute(2, 3, 0, 1), None, stride=stride, groups=c)


def swish(x):
    return x * torch.sigmoid(x)


def resize_tables(n_in: int, n_out: int):
    """ResizeBilinear(align_corners=True) interpolation tables [TF-ext]; computed in float32 like TF:
    scale = (in-1)/float(out-1); in = i*scale; lower=floor(in); upper=min(ceil(in), in-1); lerp=in-lower."""
    scale = np.float32((n_in - 1) / np.float32(n_out - 1)) if n_out > 1 else np.float32(0.0)
    i = np.arange(n_out, dtype=np.float32)
    src = (i * scale).astype(np.float32)
    lower = np.floor(src).astype(np.int64)
    upper = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
    lerp = (src - lower.astype(np.float32)).astype(np.float32)
    return lower, upper, lerp


def resize_bilinear_ac(x, out_h: int, out_w: int):
    """x NCHW.  efficientlab.py:171-172, :205-206 (align_corners=True)."""
    if x.shape[2] == out_h and x.shape[3] == out_w:
        return x            # scale == 1: lerp == 0 everywhere -> exact identity
    ylo, yhi, yl = resize_tables(x.shape[2], out_h)
    xlo, xhi, xl = resize_tables(x.shape[3], out_w)
    yl = torch.from_numpy(yl).to(x.dtype)[None, None, :, None]
    xl = torch.from_numpy(xl).to(x.dtype)[None, None, None, :]
    top_rows = x[:, :, torch.from_numpy(ylo), :]
    bot_rows = x[:, :, torch.from_numpy(yhi), :]
    xlo_t, xhi_t = torch.from_numpy(xlo), torch.from_numpy(xhi)
    tl, tr = top_rows[:, :, :, xlo_t], top_rows[:, :, :, xhi_t]
    bl, br = bot_rows[:, :, :, xlo_t], bot_rows[:, :, :, xhi_t]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return top + (bot - top) * yl


class EfficientLabOracle:
    """Functional restatement.  State is explicit: theta (flat, TF creation order), bn_state [2,n_bn]."""

    def __init__(self, arch: Optional[Arch] = None, dtype=torch.float64, dice: bool = True, l2: bool = True,
                 label_smoothing: float = 0.0, binary_iou_loss: bool = True):
        self.arch = arch or Arch()
        self.dtype = dtype
        self.dice = dice
        self.l2 = l2
        self.label_smoothing = label_smoothing
        # efficientlab.py:355-382: binary_iou_loss=True scores channel 1 only (few-shot); False (joint_train.py:307)
        # scores the flattened (H, W, C) tensors of ALL channels
        self.binary_iou_loss = binary_iou_loss
        self._l2_mask = self.arch.l2_mask(dtype)

    # -- helpers --
    def _p(self, theta, name):
        s = self.arch.by_name[name]
        return theta[s.offset:s.offset + s.size].view(s.shape)

    def _bn(self, x, theta, bn_state, new_bn, scope, training):
        """utils.py:87-134 (non-fused, biased EMA variance) / tf.layers.batch_normalization (fused:
        Bessel-corrected EMA variance) [TF-ext]; EMA: moving -= (moving - batch) * (1 - momentum)."""
        spec = self.arch.bn_by_name[scope]
        gamma = self._p(theta, scope + '/gamma')
        beta = self._p(theta, scope + '/beta')
        sl = slice(spec.offset, spec.offset + spec.channels)
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            if new_bn is not None:
                n = x.numel() // x.shape[1]
                var_ema = var * (n / (n - 1.0)) if spec.fused else var
                new_bn[0, sl] = bn_state[0, sl] - (bn_state[0, sl] - mean.detach()) * (1 - BN_MOMENTUM)
                new_bn[1, sl] = bn_state[1, sl] - (bn_state[1, sl] - var_ema.detach()) * (1 - BN_MOMENTUM)
        else:
            mean, var = bn_state[0, sl], bn_state[1, sl]
        inv = torch.rsqrt(var + BN_EPS) * gamma
        return x * inv[None, :, None, None] + (beta - mean * inv)[None, :, None, None]

    # -- forward --
    def forward(self, theta, bn_state, images_nhwc, training: bool, dc_masks=None, dropout_mask=None,
                drop_rate: float = 0.0, taps: Optional[Dict[str, torch.Tensor]] = None):
        """Returns (logits NHWC [B,H,W,2], new_bn_state).

        dc_masks: [n_dc_blocks, B] of {0,1} (binary_tensor of utils.py:157-170); None => all keep.
        dropout_mask: [B,h,w,D] of {0,1} keep mask for the final-layer dropout (efficientlab.py:161-162).
        taps: optional dict that receives intermediate activations (NHWC) for per-layer debugging.
        """
        a = self.arch
        dt = self.dtype
        x = images_nhwc.to(dt)
        # efficientlab.py:113-114
        x = (x - torch.tensor(MEAN_RGB, dtype=dt)) / torch.tensor(STDDEV_RGB, dtype=dt)
        x = x.permute(0, 3, 1, 2)
        B, _, H, W = x.shape
        new_bn = bn_state.clone() if training else None
        pre = 'efficientnet-b0/model/'

        def tap(name, t):
            if taps is not None:
                taps[name] = t.detach().permute(0, 2, 3, 1).contiguous()

        # stem: efficientnet_model.py:410-412
        x = conv2d_same(x, self._p(theta, pre + 'stem/conv2d/kernel'), stride=2)
        tap('stem.conv', x)
        x = swish(self._bn(x, theta, bn_state, new_bn, pre + 'stem/tpu_batch_normalization', training))
        tap('stem.out', x)
        endpoints = {}
        dc_i = 0
        for i, b in enumerate(a.blocks):
            sc = pre + 'blocks_%d/' % i
            inp = x
            nconv = nbn = 0
            if b.expand != 1:
                x = conv2d_same(x, self._p(theta, sc + 'conv2d/kernel'))
                tap('b%d.expand' % i, x)
                x = swish(self._bn(x, theta, bn_state, new_bn, sc + 'tpu_batch_normalization', training))
                nconv, nbn = 1, 1
            x = depthwise_same(x, self._p(theta, sc + 'depthwise_conv2d/depthwise_kernel'), b.stride)
            tap('b%d.dw' % i, x)
            bn_name = 'tpu_batch_normalization' + ('_%d' % nbn if nbn else '')
            x = swish(self._bn(x, theta, bn_state, new_bn, sc + bn_name, training))
            nbn += 1
            # SE: efficientnet_model.py:238-251
            se = x.mean(dim=(2, 3), keepdim=True)
            se = conv2d_same(se, self._p(theta, sc + 'se/conv2d/kernel'), bias=self._p(theta, sc + 'se/conv2d/bias'))
            se = swish(se)
            se = conv2d_same(se, self._p(theta, sc + 'se/conv2d_1/kernel'), bias=self._p(theta, sc + 'se/conv2d_1/bias'))
            tap('b%d.gate' % i, torch.sigmoid(se))
            x = torch.sigmoid(se) * x
            cn = 'conv2d' + ('_%d' % nconv if nconv else '')
            x = conv2d_same(x, self._p(theta, sc + cn + '/kernel'))
            tap('b%d.project' % i, x)
            bn_name = 'tpu_batch_normalization' + ('_%d' % nbn if nbn else '')
            x = self._bn(x, theta, bn_state, new_bn, sc + bn_name, training)
            if b.skip:
                if b.dc_rate > 0 and training:
                    keep = 1.0 - b.dc_rate
                    if dc_masks is not None:
                        m = dc_masks[dc_i].to(dt).view(B, 1, 1, 1)
                    else:
                        m = torch.ones(B, 1, 1, 1, dtype=dt)
                    x = (x / keep) * m
                if b.dc_rate > 0:
                    dc_i += 1
                x = x + inp
            tap('b%d.out' % i, x)
            endpoints[i] = x
        skips = {r: endpoints[a.reduction_block[r]] for r in (1, 2, 3, 4)}
        decoded = skips[4]
        # decoder: efficientlab.py:153-159
        for r in a.rsd:
            decoded = self._rsd(theta, bn_state, new_bn, decoded, skips[r], r - 1, training, tap)
        if drop_rate > 0 and training:
            # Keras Dropout [TF-ext]: keep where U >= rate, scale 1/(1-rate)
            if dropout_mask is not None:
                decoded = decoded * dropout_mask.to(dt).permute(0, 3, 1, 2) / (1.0 - drop_rate)
        tap('head.in', decoded)
        z = conv2d_same(decoded, self._p(theta, 'decode/final_layer_weights/kernel'),
                        bias=self._p(theta, 'decode/final_layer_weights/bias'))
        tap('head.logits_lowres', z)
        z = resize_bilinear_ac(z, H, W)
        return z.permute(0, 2, 3, 1).contiguous(), new_bn

    def _rsd(self, theta, bn_state, new_bn, deep, skip, idx, training, tap):
        """efficientlab.py:179-231.  conv(+bias) -> swish -> BN (efficientlab.py:185-190)."""
        sc = 'decode/decode_skip_connections_%d/' % idx
        up = resize_bilinear_ac(deep, skip.shape[2], skip.shape[3])
        tap('rsd%d.up' % idx, up)
        cat = torch.cat([up, skip], dim=1)

        def branch(t, conv, bn, dil=1):
            t = conv2d_same(t, self._p(theta, sc + conv + '/kernel'), dilation=dil, bias=self._p(theta, sc + conv + '/bias'))
            tap('rsd%d.%s' % (idx, conv), t)
            t = swish(t)
            return self._bn(t, theta, bn_state, new_bn, sc + bn, training)
        b0 = branch(cat, 'conv2d', 'batch_normalization')
        b1 = branch(cat, 'conv2d_1', 'batch_normalization_1', 2)
        b2 = cat.mean(dim=(2, 3), keepdim=True).expand(-1, -1, cat.shape[2], cat.shape[3])
        pyr = torch.cat([b0, b1, b2], dim=1)
        tap('rsd%d.pyr' % idx, pyr)
        out = branch(pyr, 'conv2d_2', 'batch_normalization_2')
        out = out + up
        tap('rsd%d.out' % idx, out)
        return out

    # -- loss: efficientlab.py:294-327, :329-396 ; regularizers.py:4-10 --
    def loss(self, theta, logits, labels):
        B = logits.shape[0]
        labels = labels.to(self.dtype)
        logp = F.log_softmax(logits, dim=-1)
        y_ce = labels
        if self.label_smoothing > 0:
            y_ce = labels * (1 - self.label_smoothing) + self.label_smoothing / labels.shape[-1]
        ce = -(y_ce * logp).sum(-1).mean()      # SUM_BY_NONZERO_WEIGHTS == mean over B*H*W rows [TF-ext]
        loss = ce
        if self.dice:
            if self.binary_iou_loss:
                p1 = torch.softmax(logits, dim=-1)[..., 1].reshape(B, -1)
                y1 = labels[..., 1].reshape(B, -1)
            else:
                p1 = torch.softmax(logits, dim=-1).reshape(B, -1)
                y1 = labels.reshape(B, -1)
            inter = (p1 * y1).sum(1)
            den = p1.sum(1) + y1.sum(1) - inter
            iou = ((inter + 1e-7) / (den + 1e-7)).mean()
            loss = loss - torch.log(2.0 * iou / (iou + 1.0))
        if self.l2:
            loss = loss + L2_COEF * 0.5 * (self._l2_mask * theta * theta).sum()
        return loss

    def loss_and_grad(self, theta, bn_state, images, labels, dc_masks=None, dropout_mask=None, drop_rate=0.0,
                      taps=None):
        th = theta.detach().clone().requires_grad_(True)
        logits, new_bn = self.forward(th, bn_state, images, True, dc_masks, dropout_mask, drop_rate, taps)
        loss = self.loss(th, logits, labels)
        (g,) = torch.autograd.grad(loss, th)
        return loss.detach(), g, new_bn, logits.detach()

    # -- predictions: efficientlab.py:174-176, :291-292 --
    def predict(self, theta, bn_state, images):
        with torch.no_grad():
            logits, _ = self.forward(theta, bn_state, images, False)
            probs = torch.softmax(logits, dim=-1)
            return (probs > 0.5).to(torch.float32), logits


class OptState:
    """Adam slot 'v' (+ beta powers) or nothing for SGD.  TF ApplyAdam [TF-ext]:
    alpha = lr*sqrt(1-b2p)/(1-b1p); m = b1*m+(1-b1)*g; v = b2*v+(1-b2)*g^2; var -= alpha*m/(sqrt(v)+eps);
    then b1p *= b1; b2p *= b2 (after all variables)."""

    def __init__(self, n, dtype, sgd=False):
        self.sgd = sgd
        self.v = torch.zeros(n, dtype=dtype)
        self.m = torch.zeros(n, dtype=dtype)
        self.b1p = ADAM_BETA1
        self.b2p = ADAM_BETA2

    def clone(self):
        o = OptState(0, self.v.dtype, self.sgd)
        o.v, o.m, o.b1p, o.b2p = self.v.clone(), self.m.clone(), self.b1p, self.b2p
        return o

    def apply(self, theta, g, lr):
        if self.sgd:
            return theta - lr * g
        alpha = lr * math.sqrt(1 - self.b2p) / (1 - self.b1p)
        self.m = ADAM_BETA1 * self.m + (1 - ADAM_BETA1) * g
        self.v = ADAM_BETA2 * self.v + (1 - ADAM_BETA2) * g * g
        theta = theta - alpha * self.m / (torch.sqrt(self.v) + ADAM_EPS)
        self.b1p *= ADAM_BETA1
        self.b2p *= ADAM_BETA2
        return theta


def iou_counts(pred_hw2: np.ndarray, label_hw2: np.ndarray) -> Tuple[int, int]:
    """reptile.py:526-549 integer part: (sum(pred & label), sum(pred | label)) on channel 1 after np.round."""
    p = np.round(pred_hw2[:, :, 1])
    l = np.round(label_hw2[:, :, 1])
    return int(np.sum(np.logical_and(p, l))), int(np.sum(np.logical_or(l, p)))


def iou_score(pred_hw2, label_hw2, epsilon=1e-7) -> float:
    i, u = iou_counts(pred_hw2, label_hw2)
    return (i + epsilon) / (u + epsilon)
