"""CPU restatement of the meta-learning steps of the reference: Reptile (`Gecko.train_step`), FOMAML
(`FOMLIS.train_step` + `FOMLIS._mini_batches`), the per-task evaluation (`Gecko._evaluate`) and the numpy list
arithmetic of the meta-update.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  PARITY UNPINNED for the float path (TF-1.15 cannot run here; the
reference has no golden vectors); the INTEGER work restated here (which task, which examples, in which order) is pinned
against the reference's own functions executed from their source text (tests/golden/sampler_sequences.json).

Reference files restated (paths into /root/reference):
  meta_learners/supervised_reptile/supervised_reptile/reptile.py:64-125    Gecko.train_step
  meta_learners/supervised_reptile/supervised_reptile/reptile.py:235-294   Gecko._evaluate (+ :482-549)
  meta_learners/supervised_reptile/supervised_reptile/reptile.py:605-663   FOMLIS.train_step, FOMLIS._mini_batches
  meta_learners/variables.py:9-55                                         interpolate / average / subtract / add / scale,
                                                                           weight_decay
  meta_learners/metaseg.py:233-343                                        task draw, mini-batches, train/test split

State semantics that matter: `_model_state` covers the TRAINABLES only
(reptile.py:34), so `import_variables(old_vars)` between the tasks of a meta-batch (reptile.py:123, :645) resets theta
but NOT the Adam slots, the beta powers or the BN moving statistics: those flow sequentially from task to task.
`_full_state` (reptile.py:35-36) covers every global variable and is what `_evaluate` saves / restores (:258, :293).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .efficientlab_oracle import EfficientLabOracle, OptState, iou_counts


# --------------------------------------------------------------------------------------------------
# integer work: meta_learners/metaseg.py, index space (an "example" is a row of the task's record pool)
# --------------------------------------------------------------------------------------------------
def sample_task(dataset: Sequence, num_shots: int):
    """metaseg.py:233-255: `random.sample(l, 1)[0]`, then the FIRST num_shots records of that task in file order
    (BinarySegmentationTask.sample, metaseg.py:214-230; the tf.data pipeline has no example-level shuffle)."""
    task = random.sample(list(dataset), 1)[0]
    n = min(num_shots, task.batch_size)        # metaseg.py:247-249
    return task, list(range(n))


def mini_batches(rows: Sequence[int], batch_size: int, num_batches: int, replacement: bool = False):
    """metaseg.py:258-302 without an augmenter."""
    rows = list(rows)
    if len(rows) == 0:
        raise ValueError("No samples to sample.")
    if replacement:
        for _ in range(num_batches):
            yield random.sample(rows, batch_size)
        return
    cur, count = [], 0
    while True:
        random.shuffle(rows)
        for r in rows:
            cur.append(r)
            if len(cur) < batch_size:
                continue
            yield cur
            cur = []
            count += 1
            if count == num_batches:
                return


def split_train_test(rows: Sequence[int], test_shots: int):
    """metaseg.py:321-343: shuffle a copy, the last test_shots rows are the test set."""
    rows = list(rows)[:]
    random.shuffle(rows)
    return rows[:-test_shots], rows[-test_shots:]


# --------------------------------------------------------------------------------------------------
# float work
# --------------------------------------------------------------------------------------------------
@dataclass
class MetaState:
    """Every global variable of the reference graph: trainables, BN moving statistics, optimizer slots."""
    theta: torch.Tensor
    bn: torch.Tensor
    opt: OptState

    def clone(self) -> "MetaState":
        return MetaState(self.theta.clone(), self.bn.clone(), self.opt.clone())


def _minimize(orc: EfficientLabOracle, st: MetaState, x, y, lr: float, dc_masks=None) -> None:
    """One `sess.run(minimize_op)`: forward, loss, backward, BN EMA update, optimizer apply (efficientlab.py:315-317)."""
    _, g, st.bn, _ = orc.loss_and_grad(st.theta, st.bn, x, y, dc_masks)
    st.theta = st.opt.apply(st.theta, g, lr)


def _batch(task, rows):
    images, labels = task.arrays()
    idx = np.asarray(rows, np.int64)
    return torch.from_numpy(images[idx]), torch.from_numpy(labels[idx])


def reptile_train_step(orc: EfficientLabOracle, st: MetaState, dataset: Sequence, num_shots: int,
                       inner_batch_size: int, inner_iters: int, replacement: bool, meta_step_size: float,
                       meta_batch_size: int, lr: Optional[float] = None, default_lr: float = 1e-3,
                       lr_scheduler: Optional[Callable[[int], float]] = None,
                       weight_decay_rate: Optional[float] = None) -> None:
    """Gecko.train_step (reptile.py:64-125), in place on `st`."""
    old = st.theta.clone()                                        # :102 export_variables (trainables)
    new_vars: List[torch.Tensor] = []
    for _ in range(meta_batch_size):                              # :104
        task, rows = sample_task(dataset, num_shots)              # :107
        for i, b in enumerate(mini_batches(rows, inner_batch_size, inner_iters, replacement)):   # :108
            x, y = _batch(task, b)
            if weight_decay_rate is not None:                     # :112-113 pre_step_op, variables.py:48-55
                st.theta = st.theta * weight_decay_rate
            # :114-121 - NB `if / if / else`: with lr given and no scheduler the minimize op runs TWICE per batch
            if lr is not None:
                _minimize(orc, st, x, y, lr)
            if lr_scheduler is not None:
                _minimize(orc, st, x, y, lr_scheduler(i))
            else:
                _minimize(orc, st, x, y, default_lr)
        new_vars.append(st.theta.clone())                         # :122
        st.theta = old.clone()                                    # :123 (optimizer slots / BN statistics keep flowing)
    mean = torch.stack(new_vars).mean(0)                          # :124 average_vars, variables.py:16-23
    st.theta = old + (mean - old) * meta_step_size                # :125 interpolate_vars, variables.py:9-13


def fomaml_mini_batches(rows, inner_batch_size, inner_iters, replacement, tail_shots: Optional[int]):
    """FOMLIS._mini_batches (reptile.py:649-663) without replacement-sampled train/val."""
    if tail_shots is None:
        yield from mini_batches(rows, inner_batch_size, inner_iters, replacement)
        return
    train, tail = split_train_test(rows, tail_shots)
    yield from mini_batches(train, inner_batch_size, inner_iters - 1, replacement)
    yield tail


def fomaml_train_step(orc: EfficientLabOracle, st: MetaState, dataset: Sequence, num_shots: int,
                      inner_batch_size: int, inner_iters: int, replacement: bool, meta_step_size: float,
                      meta_batch_size: int, tail_shots: Optional[int] = None, lr: Optional[float] = None,
                      default_lr: float = 1e-3, weight_decay_rate: Optional[float] = None) -> None:
    """FOMLIS.train_step (reptile.py:605-647), in place on `st`."""
    old = st.theta.clone()
    updates: List[torch.Tensor] = []
    for _ in range(meta_batch_size):
        task, rows = sample_task(dataset, num_shots)
        last_backup = None
        for j, b in enumerate(fomaml_mini_batches(rows, inner_batch_size, inner_iters, replacement, tail_shots)):
            x, y = _batch(task, b)
            if j == inner_iters - 1:
                last_backup = st.theta.clone()                    # :635-636
            if weight_decay_rate is not None:
                st.theta = st.theta * weight_decay_rate
            _minimize(orc, st, x, y, lr if lr is not None else default_lr)     # :639-643 (a proper if / else)
        updates.append(st.theta - last_backup)                    # :644 subtract_vars
        st.theta = old.clone()                                    # :645
    update = torch.stack(updates).mean(0)                         # :646 average_vars
    st.theta = old + update * meta_step_size                      # :647 add_vars(old, scale_vars(update, eps))


def evaluate_task(orc: EfficientLabOracle, st: MetaState, task, num_shots: int, test_shots: int,
                  inner_batch_size: int, inner_iters: int, replacement: bool, lr: Optional[float] = None,
                  default_lr: float = 1e-3) -> Tuple[float, List[Tuple[int, int]]]:
    """One task of Gecko.evaluate (reptile.py:195-204) + Gecko._evaluate (:235-294), transductive prediction.
    `st` is left untouched (the reference restores `_full_state`).  Returns (mean IoU, per-image (inter, union))."""
    _, rows = sample_task([task], num_shots + test_shots)
    train, test = split_train_test(rows, test_shots)
    w = st.clone()                                                # :258 _full_state.export_variables()
    for b in mini_batches(train, inner_batch_size, inner_iters, replacement):
        x, y = _batch(task, b)
        _minimize(orc, w, x, y, lr if lr is not None else default_lr)
    xq, _ = _batch(task, test)
    pred, _ = orc.predict(w.theta, w.bn, xq)                      # :503-506, is_training_ph False
    _, labels = task.arrays()
    counts = [iou_counts(pred[j].numpy(), labels[test[j]]) for j in range(len(test))]
    ious = [(i + 1e-7) / (u + 1e-7) for i, u in counts]           # :549
    return float(np.nanmean(ious)), counts                        # :290-291


def state_from_flat(theta_tf_order, bn_2xn, adam_v_tf_order=None, beta1_power: float = 0.0,
                    beta2_power: float = 0.999, sgd: bool = False, dtype=torch.float64) -> MetaState:
    """A MetaState from flat vectors in tf.trainable_variables() order (what the engine under test exports)."""
    theta = torch.as_tensor(theta_tf_order).detach().cpu().to(dtype).clone()
    bn = torch.as_tensor(bn_2xn).detach().cpu().to(dtype).clone()
    opt = OptState(theta.numel(), dtype, sgd=sgd)
    if adam_v_tf_order is not None:
        opt.v = torch.as_tensor(adam_v_tf_order).detach().cpu().to(dtype).clone()
    opt.b1p, opt.b2p = float(beta1_power), float(beta2_power)
    return MetaState(theta, bn, opt)
