"""ctypes binding of libmliis_b200.so (include/mliis_b200.h).

There is no CPU fallback: if the library is missing this module raises at import of the symbol table, and every
compute entry point fails with MLIIS_ERR_DEVICE on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, NamedTuple, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmliis_b200.so")

MLIIS_OK, MLIIS_ERR_ARG, MLIIS_ERR_CUDA, MLIIS_ERR_DEVICE, MLIIS_ERR_STATE = 0, -1, -2, -3, -4
OPT_ADAM, OPT_SGD = 0, 1
LOSS_DICE, LOSS_L2 = 1, 2
GEMM_FP32, GEMM_TF32, GEMM_TF32X3 = 0, 1, 2
STATE_TRAINABLES, STATE_BN, STATE_OPT, STATE_ALL = 1, 2, 4, 7


class Config(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("max_batch", C.c_int32), ("n_slots", C.c_int32),
                ("optimizer", C.c_int32), ("loss_flags", C.c_int32), ("gemm_mode", C.c_int32),
                ("label_smoothing", C.c_float), ("final_dropout_rate", C.c_float), ("rsd", C.c_int32 * 4),
                ("n_classes", C.c_int32)]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_int64), ("size", C.c_int64), ("ndim", C.c_int32),
                ("shape", C.c_int32 * 4), ("l2", C.c_int32)]


class BnInfo(C.Structure):
    _fields_ = [("scope", C.c_char_p), ("channels", C.c_int32), ("offset", C.c_int32), ("fused", C.c_int32)]


class StepArgs(C.Structure):
    _fields_ = [("dev_images", C.c_void_p), ("dev_labels", C.c_void_p), ("dev_index", C.c_void_p),
                ("batch", C.c_int32), ("lr", C.c_float), ("pre_decay_rate", C.c_float),
                ("dev_dc_mask", C.c_void_p), ("dev_drop_mask", C.c_void_p), ("seed", C.c_uint64),
                ("dev_loss_out", C.c_void_p), ("dev_seed", C.c_void_p)]


class TaskArgs(C.Structure):
    _fields_ = [("dev_init_state", C.c_void_p), ("dev_images", C.c_void_p), ("dev_labels", C.c_void_p),
                ("dev_batch_index", C.c_void_p), ("dev_lr", C.c_void_p), ("n_steps", C.c_int32),
                ("batch", C.c_int32), ("dev_query_index", C.c_void_p), ("n_query", C.c_int32),
                ("dev_dc_mask", C.c_void_p), ("seed", C.c_uint64), ("pre_decay_rate", C.c_float),
                ("dev_inter_out", C.c_void_p), ("dev_union_out", C.c_void_p), ("dev_loss_out", C.c_void_p),
                ("dev_seed", C.c_void_p), ("n_group", C.c_int32), ("group_stride_bytes", C.c_int64)]


# every symbol include/mliis_b200.h declares: (name, restype, argtypes)
_VP, _I32, _I64, _F, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
SYMBOLS = [
    ("mliis_last_error", C.c_char_p, []),
    ("mliis_version", C.c_char_p, []),
    ("mliis_ctx_create", C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_VP)]),
    ("mliis_ctx_destroy", C.c_int, [_VP]),
    ("mliis_num_params", _I64, [_VP]),
    ("mliis_num_param_tensors", _I32, [_VP]),
    ("mliis_num_bn_layers", _I32, [_VP]),
    ("mliis_num_bn_channels", _I32, [_VP]),
    ("mliis_num_dc_blocks", _I32, [_VP]),
    ("mliis_param_table", C.c_int, [_VP, C.POINTER(ParamInfo), _I32]),
    ("mliis_bn_table", C.c_int, [_VP, C.POINTER(BnInfo), _I32]),
    ("mliis_workspace_bytes", _I64, [_VP]),
    ("mliis_state_floats", _I64, [_VP]),
    ("mliis_theta_floats", _I64, [_VP]),
    ("mliis_slot_bind", C.c_int, [_VP, _I32, _VP, _VP]),
    ("mliis_state_copy", C.c_int, [_VP, _VP, _VP, _I32, _VP]),
    ("mliis_train_step", C.c_int, [_VP, _I32, C.POINTER(StepArgs), _VP]),
    ("mliis_forward", C.c_int, [_VP, _I32, _VP, _VP, _I32, _I32, _VP, _VP, _U64, _VP, _VP]),
    ("mliis_loss_backward", C.c_int, [_VP, _I32, _VP, _VP, _I32, _VP, _VP, _VP]),
    ("mliis_optimizer_step", C.c_int, [_VP, _I32, _F, _F, _VP]),
    ("mliis_set_grads", C.c_int, [_VP, _I32, _VP, _VP]),
    ("mliis_set_class_ids", C.c_int, [_VP, _I32, _VP]),
    ("mliis_predict_classes", C.c_int, [_VP, _I32, _VP, _VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    ("mliis_predict", C.c_int, [_VP, _I32, _VP, _VP, _VP, _I32, _VP, _VP, _VP, _VP, _VP]),
    ("mliis_adapt_eval_task", C.c_int, [_VP, _I32, C.POINTER(TaskArgs), _VP]),
    ("mliis_task_graph_capture", C.c_int, [_VP, _I32, C.POINTER(TaskArgs), _VP]),
    ("mliis_task_graph_launch", C.c_int, [_VP, _I32, _VP]),
    ("mliis_launch_count", C.c_uint64, []),
    ("mliis_crc32c", C.c_uint32, [_VP, C.c_uint64, C.c_uint32]),
    ("mliis_delta_accumulate", C.c_int, [_VP, _VP, _VP, _VP, _I32, _VP]),
    ("mliis_meta_apply", C.c_int, [_VP, _VP, _VP, _F, _VP]),
    ("mliis_meta_buffer_floats", _I64, [_VP]),
    ("mliis_comm_unique_id", C.c_int, [_VP]),
    ("mliis_comm_init", C.c_int, [_VP, _VP, _I32, _I32]),
    ("mliis_comm_destroy", C.c_int, [_VP]),
    ("mliis_allreduce_delta", C.c_int, [_VP, _VP, _I64, _VP]),
    ("mliis_meta_reduce", C.c_int, [_VP, _VP, _VP, _I64, _I32, _I32, _VP]),
    ("mliis_meta_finish", C.c_int, [_VP, _VP, _VP, _F, _I32, _I32, _VP]),
    ("mliis_dwconv_fwd", C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP, _VP, _VP]),
    ("mliis_gemm_nn", C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_conv3x3_fwd", C.c_int, [_VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_tc_prep_weights", C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_tc_conv", C.c_int, [_VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_tc_project_conv", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_tc_wgrad", C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_bilinear_fwd", C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_adam_step", C.c_int, [_VP, _VP, _VP, _I64, _I64, _F, _F, _F, _VP]),
    ("mliis_kernel_group", C.c_int, [_I32, _I64]),
    ("mliis_kernel_scratch_floats", _I64, [_I32, _I32, _I32, _I32]),
    ("mliis_dwconv_bwd", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_bn_stats_fwd", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _VP]),
    ("mliis_bn_swish_bwd", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _VP]),
    ("mliis_se_fwd", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_softmax_ce_iou", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _F, _VP]),
    ("mliis_tc_prep_weights_sub", C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    ("mliis_rsd_conv2_fwd", C.c_int, [_VP, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                      _VP]),
    ("mliis_tc_peak_tf32", C.c_int, [_I32, C.POINTER(C.c_double), _VP]),
    ("mliis_tc_mma_rate", C.c_int, [_I32, _I32, _I32, _I32, C.POINTER(C.c_double), _VP]),
    ("mliis_debug_buffer", C.c_int, [_VP, _I32, C.c_char_p, C.POINTER(_VP), C.POINTER(_I64), C.POINTER(_I32),
                                     C.POINTER(_I32)]),
]


class MliisError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("mliis_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m mliis_b200.build` (nvcc, sm_100a). "
                              "mliis_b200 has no CPU / PyTorch fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)      # AttributeError if the ABI and the header disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != MLIIS_OK:
        raise MliisError(rc, (lib().mliis_last_error() or b"").decode())


class Param(NamedTuple):
    name: str
    offset: int
    size: int
    shape: Tuple[int, ...]
    l2: bool


class BnLayer(NamedTuple):
    scope: str
    channels: int
    offset: int
    fused: bool


def make_config(image_size=224, max_batch=8, n_slots=1, optimizer=OPT_ADAM, loss_flags=LOSS_DICE | LOSS_L2,
                gemm_mode=GEMM_FP32, label_smoothing=0.0, final_dropout_rate=0.0, rsd=(2, 4), n_classes=1) -> Config:
    cfg = Config()
    cfg.image_size, cfg.max_batch, cfg.n_slots = image_size, max_batch, n_slots
    cfg.optimizer, cfg.loss_flags, cfg.gemm_mode = optimizer, loss_flags, gemm_mode
    cfg.label_smoothing, cfg.final_dropout_rate = label_smoothing, final_dropout_rate
    cfg.n_classes = n_classes
    r = list(rsd or ())[:4]
    for i in range(4):
        cfg.rsd[i] = r[i] if i < len(r) else 0
    return cfg


class Context:
    """Owns a mliis_ctx*.  device=-1 gives a table-only context (no CUDA call; works without a GPU)."""

    def __init__(self, cfg: Config, device: int):
        self.cfg = cfg
        self.device = device
        self._h = _VP()
        check(lib().mliis_ctx_create(C.byref(cfg), device, C.byref(self._h)))
        l = lib()
        self.n_params = l.mliis_num_params(self._h)
        self.n_theta = l.mliis_theta_floats(self._h)
        self.n_bn = l.mliis_num_bn_channels(self._h)
        self.n_dc = l.mliis_num_dc_blocks(self._h)
        self.state_floats = l.mliis_state_floats(self._h)
        self.workspace_bytes = l.mliis_workspace_bytes(self._h)
        n = l.mliis_num_param_tensors(self._h)
        arr = (ParamInfo * n)()
        check(l.mliis_param_table(self._h, arr, n))
        self.params: List[Param] = [Param(a.name.decode(), a.offset, a.size, tuple(a.shape[:a.ndim]), bool(a.l2))
                                    for a in arr]
        nb = l.mliis_num_bn_layers(self._h)
        barr = (BnInfo * nb)()
        check(l.mliis_bn_table(self._h, barr, nb))
        self.bns: List[BnLayer] = [BnLayer(b.scope.decode(), b.channels, b.offset, bool(b.fused)) for b in barr]

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().mliis_ctx_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
