"""Variable-list arithmetic and VariableState (mirror of /root/reference/meta_learners/variables.py).

The list arithmetic (:9-45) is kept for API compatibility (numpy, host).  On the device fast path the same
operations are the fused kernels mliis_delta_accumulate / mliis_meta_apply on the flat parameter buffer.
``VariableState`` (:58-80) keeps export/import through host numpy for drop-in use and adds
``snapshot()/restore()`` which stay on the device.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from .session import WeightDecayOp


def interpolate_vars(old_vars, new_vars, epsilon):
    """old + epsilon * (new - old)  (variables.py:9-13)."""
    return add_vars(old_vars, scale_vars(subtract_vars(new_vars, old_vars), epsilon))


def average_vars(var_seqs):
    """Element-wise mean over a sequence of variable lists (variables.py:16-23)."""
    return [np.mean(variables, axis=0) for variables in zip(*var_seqs)]


def subtract_vars(var_seq_1, var_seq_2):
    return [a - b for a, b in zip(var_seq_1, var_seq_2)]


def add_vars(var_seq_1, var_seq_2):
    return [a + b for a, b in zip(var_seq_1, var_seq_2)]


def scale_vars(var_seq, scale):
    return [v * scale for v in var_seq]


def weight_decay(rate, variables=None):
    """An op that multiplies every trainable variable by `rate` (variables.py:48-55)."""
    if variables is not None:
        raise NotImplementedError("weight decay over a subset of variables")
    return WeightDecayOp(rate)


_SEGMENTS = ("theta", "moving_mean", "moving_variance", "adam_v", "beta1_power", "beta2_power")


class VariableState:
    """Save / restore a set of variables of the model bound to `session`."""

    def __init__(self, session, variables: Sequence):
        self._session = session
        self._variables = list(variables)

    def _views(self):
        eng = self._session.model.engine()
        bn = eng.bn_state(0)
        return {"theta": eng.theta(0), "moving_mean": bn[0], "moving_variance": bn[1], "adam_v": eng.adam_v(0),
                "beta1_power": eng.powers(0)[0:1], "beta2_power": eng.powers(0)[1:2]}

    def export_variables(self) -> List[np.ndarray]:
        """Host copies of the variables, in collection order (variables.py:70-74)."""
        host = {k: v.detach().cpu().numpy() for k, v in self._views().items()}
        out = []
        for v in self._variables:
            a = host[v.segment][v.offset:v.offset + v.size]
            out.append(a.reshape(v.shape).copy() if v.shape else np.float32(a[0]))
        return out

    def import_variables(self, values) -> None:
        """Assign host values back to the variables (variables.py:76-80)."""
        views = self._views()
        staged = {k: views[k].detach().cpu().numpy().copy() for k in {v.segment for v in self._variables}}
        for v, val in zip(self._variables, values):
            staged[v.segment][v.offset:v.offset + v.size] = np.asarray(val, np.float32).reshape(-1)
        for k, a in staged.items():
            views[k].copy_(torch.from_numpy(a).to(views[k].device))

    # ---- device-resident variants used by the fast path ----
    def snapshot(self) -> torch.Tensor:
        return self._session.model.engine().states[0].clone()

    def restore(self, snap: torch.Tensor, what: int = 7) -> None:
        eng = self._session.model.engine()
        eng.copy_state(eng.states[0], snap, what)
