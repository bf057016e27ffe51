"""A synthetic stand-in for the missing EfficientLab-6-3_FOMAML-star checkpoint (the tarball
is absent from the reference mount), produced ON THE ENGINE so that benchmarks, smoke() and the parity tests start
from a state that actually segments: reference initialisers -> `steps` Adam steps over a few synthetic
"meta-training" tasks -> BN moving statistics (momentum 0.99: they lag hundreds of steps) replaced by the batch
statistics of one image per task.  The checkpoint keeps its optimizer slots, like a `tf.train.Saver` bundle of every
global variable would (reptile.py:35-36, run_metasegnet.py:131-133).
"""
from __future__ import annotations

import numpy as np
import torch

from .init import initial_bn_state, initial_variables
from .synthetic import make_task_arrays, parse_records

BN_MOMENTUM = 0.99          # efficientnet_builder.py:137


def synthetic_checkpoint(eng, steps: int = 100, lr: float = 1e-3, n_tasks: int = 8, per_task: int = 6,
                         batch: int = 8, first_task_id: int = 100000, seed: int = 0, slot: int = 0) -> torch.Tensor:
    """Leaves the state in `slot` of `eng` and returns a clone of it (mliis_state_floats() floats)."""
    size = eng.image_size
    eng.init_state(slot, initial_variables(eng.ctx.params, seed), *initial_bn_state(eng.n_bn))
    pools = [parse_records(*make_task_arrays(first_task_id + t, per_task, size)) for t in range(n_tasks)]
    x = torch.from_numpy(np.concatenate([p[0] for p in pools])).to(eng.device)
    y = torch.from_numpy(np.concatenate([p[1] for p in pools])).to(eng.device)
    rng = np.random.default_rng(seed)
    B = min(batch, eng.max_batch)
    for _ in range(steps):
        idx = torch.from_numpy(rng.integers(0, x.shape[0], B).astype(np.int32)).to(eng.device)
        eng.train_step(slot, x, y, lr, index=idx)
    b0 = eng.bn_state(slot).clone()
    ridx = torch.arange(0, n_tasks * per_task, per_task, dtype=torch.int32)[:eng.max_batch].to(eng.device)
    eng.forward(slot, x, True, index=ridx, want_logits=False)           # one EMA update towards the batch statistics
    torch.cuda.synchronize()
    eng.bn_state(slot).copy_(b0 + (eng.bn_state(slot) - b0) / (1.0 - BN_MOMENTUM))   # setup-time plumbing, untimed
    torch.cuda.synchronize()
    return eng.states[slot].clone()


def pretraining_pool(image_size: int, n_tasks: int = 8, per_task: int = 6, first_task_id: int = 100000):
    """The (images, labels) host arrays the checkpoint was trained on (for checks that need a segmenting input)."""
    pools = [parse_records(*make_task_arrays(first_task_id + t, per_task, image_size)) for t in range(n_tasks)]
    return np.concatenate([p[0] for p in pools]), np.concatenate([p[1] for p in pools])
