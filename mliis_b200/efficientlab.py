"""EfficientLab model object: the operator surface the meta-learner consumes, backed by the B200 engine.

Mirror of the attribute contract of /root/reference/models/efficientlab.py (:42-61, :94-108, :176, :300,
:313-317, :398): ``input_ph, label_ph, is_training_ph, lr_ph, final_layer_dropout_rate_ph, minimize_op,
predictions, loss, variables_initialized, feature_extractor_name, final_layer_scope, restore_model``.  The
"placeholders" and "ops" are opaque handles dispatched by ``session.Session.run``; no graph is built - the
network is the fixed plan compiled into libmliis_b200.so (csrc/plan.cpp).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np

from . import native as N

FINAL_LAYER_WEIGHTS_NAME = "final_layer_weights"     # efficientlab.py:17
FEATURE_DECODER_SCOPE_NAME = "decode"                # efficientlab.py:18
_GEMM_MODES = {"fp32": N.GEMM_FP32, "tf32": N.GEMM_TF32, "tf32x3": N.GEMM_TF32X3}


class Handle:
    """An opaque graph handle (placeholder / op / tensor) with a TF-like name."""

    def __init__(self, name: str, kind: str, default=None):
        self.name, self.kind, self.default = name, kind, default

    def __repr__(self):
        return "<%s %s>" % (self.kind, self.name)


class Variable:
    """One global variable of the reference graph, resolved to a slice of the engine's state vector."""

    def __init__(self, name: str, shape, segment: str, offset: int, trainable: bool):
        self.name, self.shape, self.segment, self.offset, self.trainable = name, tuple(shape), segment, offset, trainable
        self.size = int(np.prod(shape)) if len(shape) else 1


class EfficientLab:
    def __init__(self, images=None, labels=None, is_training: bool = True, n_classes=1, n_rows=224, n_cols=224,
                 spatial_pyramid_pooling: bool = False, skip_decoding: bool = False,
                 feature_extractor_name: str = "efficientnet-b0", l2: bool = True, l1: bool = False,
                 darc1: bool = False, final_layer_dropout_rate: Optional[float] = 0.2, dice: bool = True,
                 optimizer=None, rsd: Optional[List[int]] = [2], disable_lsd_residual_connections: bool = False,
                 seperate_background_channel: bool = True, binary_iou_loss: bool = True, **optim_kwargs):
        if images is not None or labels is not None:
            raise NotImplementedError("there is no TF input graph here: feed batches through mliis_b200.joint_train "
                                      "(joint_train.py path, SURVEY 8f-1)")
        if feature_extractor_name != "efficientnet-b0":
            raise NotImplementedError("only efficientnet-b0 (EfficientLab-6-3) is built; got %s" % feature_extractor_name)
        for flag, name in ((spatial_pyramid_pooling, "spatial_pyramid_pooling"), (skip_decoding, "skip_decoding"),
                           (l1, "l1"), (darc1, "darc1"), (disable_lsd_residual_connections,
                                                          "disable_lsd_residual_connections")):
            if flag:
                raise NotImplementedError("%s is not enabled by run.sh / BASELINE configs and is not built" % name)
        # two heads are built: the few-shot binary head (n_classes=1, binary_iou_loss=True) and the joint-training
        # head of joint_train.py:307 (n_classes=N, seperate_background_channel=True, binary_iou_loss=False: N+1
        # channels, multi-class soft IoU, sparse labels on the device)
        if not seperate_background_channel:
            raise NotImplementedError("a head without the background channel is not built")
        if (n_classes == 1) != bool(binary_iou_loss):
            raise NotImplementedError("built heads: n_classes=1 with binary_iou_loss=True, or n_classes>1 with "
                                      "binary_iou_loss=False")
        self.n_classes = int(n_classes)
        self.binary_iou_loss = bool(binary_iou_loss)
        if n_rows != n_cols:
            raise ValueError("square images only")
        if not rsd:
            raise NotImplementedError("a model without RSD modules is not built (run.sh uses --rsd 2 4)")
        if optimizer is None:
            optimizer = "adam"                       # DEFAULT_OPTIMIZER, efficientlab.py:16
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' (AdamOptimizer beta1=0) or 'sgd'")
        print("Using optimizer {}".format(optimizer))
        self.optimizer_name = optimizer
        self.n_input_channels = 3
        self.n_input_rows, self.n_input_cols = n_rows, n_cols
        self.n_output_channels = n_classes + 1
        self.l2, self.l1, self.darc1, self.dice = l2, l1, darc1, dice
        self.rsd = list(rsd)
        self.feature_extractor_name = feature_extractor_name
        self.aspp_dimension, self.max_block_num = 112, 10
        self.feature_decoder_name = FEATURE_DECODER_SCOPE_NAME
        self.final_layer_scope = self.feature_decoder_name + "/" + FINAL_LAYER_WEIGHTS_NAME
        self.learning_rate = float(optim_kwargs.get("learning_rate", 1e-3))
        self.label_smoothing = float(optim_kwargs.get("label_smoothing", 0.0))
        print("Label smoothing epsilon: {}".format(self.label_smoothing))
        print("Defining optimizer with default learning rate: {}".format(self.learning_rate))
        self.final_layer_dropout_rate = float(final_layer_dropout_rate or 0.0)
        self.gemm_mode = _GEMM_MODES[optim_kwargs.get("gemm_mode", "fp32")]
        self.task_slots = int(optim_kwargs.get("task_slots", 8))
        self.max_batch = int(optim_kwargs.get("max_batch", 8))

        self.input_ph = Handle("X", "placeholder")
        self.label_ph = Handle("Y", "placeholder")
        self.is_training_ph = Handle("is_training", "placeholder", default=is_training)
        self.lr_ph = Handle("learning_rate", "placeholder", default=self.learning_rate)
        if self.final_layer_dropout_rate > 0:
            print("Using dropout at final layer with drop rate {}".format(self.final_layer_dropout_rate))
            self.final_layer_dropout_rate_ph = Handle("final_layer_dropout_rate", "placeholder",
                                                      default=self.final_layer_dropout_rate)
        else:
            self.final_layer_dropout_rate_ph = self.final_layer_dropout_rate
        self.minimize_op = Handle("minimize", "op")
        self.predictions = Handle("predictions", "tensor")
        self.probabilities = Handle("probabilities", "tensor")
        self.logits = Handle("logits", "tensor")
        self.loss = Handle("loss", "tensor")
        self.variables_initialized = False

        # variable tables come from the library (no CUDA needed)
        cfg = self._config(1)
        tab = N.Context(cfg, -1)
        self.params = tab.params
        self.bns = tab.bns
        self.n_theta, self.n_bn, self.n_params = tab.n_theta, tab.n_bn, tab.n_params
        tab.close()
        self._engine = None
        self._pending_state = None
        print("final feature tensor: decode/decode_skip_connections_%d (%d channels)" % (min(self.rsd) - 1, 112))

    # ---- engine ownership ----
    def _config(self, n_slots: int):
        flags = (N.LOSS_DICE if self.dice else 0) | (N.LOSS_L2 if self.l2 else 0)
        return N.make_config(self.n_input_rows, self.max_batch, n_slots,
                             N.OPT_SGD if self.optimizer_name == "sgd" else N.OPT_ADAM, flags, self.gemm_mode,
                             self.label_smoothing, self.final_layer_dropout_rate, self.rsd, self.n_classes)

    def engine(self, device: Optional[int] = None):
        """The CUDA engine (created on first use; raises without an sm_100 GPU - there is no fallback)."""
        if self._engine is None:
            from .engine import Engine
            if device is None:
                device = int(os.environ.get("LOCAL_RANK", "0"))
            self._engine = Engine(image_size=self.n_input_rows, max_batch=self.max_batch, n_slots=self.task_slots,
                                  sgd=self.optimizer_name == "sgd", dice=self.dice, l2=self.l2,
                                  label_smoothing=self.label_smoothing,
                                  final_dropout_rate=self.final_layer_dropout_rate, rsd=self.rsd,
                                  gemm_mode=self.gemm_mode, device=device, n_classes=self.n_classes)
        return self._engine

    # ---- variable collections (tf.trainable_variables() / GLOBAL_VARIABLES order) ----
    def trainable_variables(self) -> List[Variable]:
        return [Variable(p.name, p.shape, "theta", p.offset, True) for p in self.params]

    def global_variables(self) -> List[Variable]:
        """Trainables interleaved with BN moving statistics in creation order, then the optimizer's slots
        (beta powers, per-variable Adam `v`; Adam `m` is dead for beta1=0 and is not stored)."""
        out: List[Variable] = []
        bn_iter = iter(self.bns)
        for p in self.params:
            out.append(Variable(p.name, p.shape, "theta", p.offset, True))
            if p.name.endswith("/beta"):
                b = next(bn_iter)
                out.append(Variable(b.scope + "/moving_mean", (b.channels,), "moving_mean", b.offset, False))
                out.append(Variable(b.scope + "/moving_variance", (b.channels,), "moving_variance", b.offset, False))
        if self.optimizer_name == "adam":
            out.append(Variable("beta1_power", (), "beta1_power", 0, False))
            out.append(Variable("beta2_power", (), "beta2_power", 0, False))
            for p in self.params:
                out.append(Variable(p.name + "/Adam_1", p.shape, "adam_v", p.offset, False))
        return out

    def initialize(self, seed: int = 0) -> None:
        """tf.global_variables_initializer(): reference initialisers (mliis_b200/init.py), optimizer slots zero."""
        from .init import initial_bn_state, initial_variables
        eng = self.engine()
        mm, mv = initial_bn_state(self.n_bn)
        eng.init_state(0, initial_variables(self.params, seed), mm, mv)
        self.variables_initialized = True

    def restore_model(self, sess, ckpt_dir, enable_ema=False, export_ckpt=None, filter_to_scopes=None,
                      filter_out_scope=None, convert_ckpt_to_rel_path: bool = False):
        """efficientlab.py:398-443: (re)initialise, then restore the variables selected by the scope filters."""
        assert isinstance(filter_to_scopes, list) or filter_to_scopes is None
        assert isinstance(filter_out_scope, str) or filter_out_scope is None
        if enable_ema or export_ckpt:
            raise NotImplementedError("EMA restore / re-export is not part of the hot path")
        from .checkpoint import restore_into_engine
        self.initialize()

        def keep(name: str) -> bool:
            if filter_out_scope is not None and name.startswith(filter_out_scope):
                return False
            if filter_to_scopes is not None:
                return any(name.startswith(x) for x in filter_to_scopes)
            return True
        n = restore_into_engine(self, ckpt_dir, keep=keep, strict=False, warn_missing=True)
        self.variables_initialized = True
        print("Variables initialized")
        print("{} variables restored".format(n))
