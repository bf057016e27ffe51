"""Random initialisation of the EfficientLab variables (used when no checkpoint is restored).

Mirrors the reference initialisers: ``conv_kernel_initializer`` N(0, 2/fan_out) for every EfficientNet kernel,
SE kernel and the final layer (models/efficientnet/efficientnet_model.py:61-82; efficientlab.py:102-103,:166),
tf.layers defaults (glorot_uniform kernel, zero bias) for the decoder convs (efficientlab.py:186-188), BN
gamma=1 / beta=0 / moving_mean=0 / moving_variance=1.  The random streams are numpy's, not TF's.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


def initial_variables(params: Sequence, seed: int = 0) -> List[np.ndarray]:
    """params: native.Param list (tf.trainable_variables() order) -> list of float32 arrays."""
    rng = np.random.default_rng(seed)
    out = []
    for p in params:
        name, shape = p.name, tuple(p.shape)
        if name.endswith("/gamma"):
            v = np.ones(shape, np.float32)
        elif name.endswith("/beta") or name.endswith("/bias"):
            v = np.zeros(shape, np.float32)
        elif name.startswith("decode/decode_skip_connections"):
            kh, kw, ci, co = shape
            lim = math.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
            v = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            kh, kw, _, co = shape
            v = (rng.standard_normal(size=shape) * math.sqrt(2.0 / (kh * kw * co))).astype(np.float32)
        out.append(v)
    return out


def initial_bn_state(n_bn: int) -> Tuple[np.ndarray, np.ndarray]:
    return np.zeros(n_bn, np.float32), np.ones(n_bn, np.float32)
