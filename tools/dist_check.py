#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL):
  1. meta-test: tasks sharded over ranks give the same per-task IoUs as one rank evaluating all tasks;
  2. FOMAML meta-step with --sgd: sharded tasks + one NCCL all-reduce of the summed deltas == the single-rank step.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os
import random
import sys
from functools import partial

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from mliis_b200 import reptile
from mliis_b200.efficientlab import EfficientLab
from mliis_b200.reptile import FOMLIS, Gecko
from mliis_b200.session import Session
from mliis_b200.synthetic import SyntheticSegmentationTask

SIZE = 64


def model(opt):
    m = EfficientLab(rsd=[2, 4], l2=True, dice=True, final_layer_dropout_rate=0.0, n_rows=SIZE, n_cols=SIZE,
                     learning_rate=1e-3, label_smoothing=0.0, optimizer=opt, task_slots=3, gemm_mode="tf32x3")
    m.initialize(seed=0)
    return m


tasks = [SyntheticSegmentationTask(500 + i, 10, SIZE) for i in range(8)]
kw = dict(num_classes=1, num_shots=5, inner_batch_size=8, inner_iters=3, replacement=False, eval_all_tasks=True)

# ---- 1. sharded meta-test ----
m = model("adam")
sess = Session(m)
kw_eval = dict(kw, is_training_ph=m.is_training_ph, lr_ph=m.lr_ph)
random.seed(7)
g = Gecko(sess, transductive=True)
mean_sharded, map_sharded = g.evaluate(tasks, m.input_ph, m.label_ph, m.minimize_op, m.predictions, **kw_eval)
# the same evaluation on every rank alone (world forced to 1)
real_dist = reptile._dist
reptile._dist = lambda: (0, 1)
import mliis_b200.runner as runner_mod
real_gather = runner_mod.gather_owned
runner_mod.gather_owned = lambda v, device=None: v
random.seed(7)
mean_single, map_single = Gecko(sess, transductive=True).evaluate(tasks, m.input_ph, m.label_ph, m.minimize_op,
                                                                  m.predictions, **kw_eval)
reptile._dist, runner_mod.gather_owned = real_dist, real_gather
assert map_sharded.keys() == map_single.keys()
err = max(abs(map_sharded[k] - map_single[k]) for k in map_single)
assert err < 1e-12, err
print("[rank %d] sharded meta-test == single-rank meta-test (%d tasks, max |dIoU| = %.1e)" % (rank, len(tasks), err))

# ---- 2. FOMAML meta-step, SGD ----
res = []
for sharded in (True, False):
    m = model("sgd")
    sess = Session(m)
    random.seed(3)
    if not sharded:
        reptile._dist = lambda: (0, 1)
        real_ar = runner_mod.allreduce_meta
        runner_mod.allreduce_meta = lambda d, b=None: None
    learner = FOMLIS(sess, train_shots=10, tail_shots=5)
    learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, num_classes=1, num_shots=10, inner_batch_size=4,
                       inner_iters=3, replacement=False, meta_step_size=0.5, meta_batch_size=4, lr_ph=m.lr_ph, lr=None)
    if not sharded:
        reptile._dist = real_dist
        runner_mod.allreduce_meta = real_ar
    eng = m.engine()
    res.append(eng.tf_order_vector(eng.theta(0)).double().clone())
rel = ((res[0] - res[1]).norm() / res[1].norm()).item()
assert rel < 1e-6, rel
# theta is replicated: bit-identical across ranks after the all-reduce
t = res[0].clone()
dist.broadcast(t, 0)
assert torch.equal(t, res[0])
print("[rank %d] sharded FOMAML step + NCCL all-reduce == single-rank step (rel L2 %.1e); theta identical across ranks" % (rank, rel))

# ---- 3. slot-parallel meta-steps where some ranks own NO task of the meta-batch (8 GPUs on FOMAML's meta-batch of 5) ----
# meta-batch 1 on >= 2 ranks: rank 0 adapts the task on its training slots, every other rank contributes zeros to the ONE
# all-reduce and applies the same update; then a meta-batch of world + 1 (rank 0: two tasks, the others one each).
m = model("adam")
sess = Session(m)
random.seed(11)
learner = Gecko(sess, meta_task_slots=3)
for mb in (1, world + 1, 1):
    learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, num_classes=1, num_shots=5, inner_batch_size=4,
                       inner_iters=2, replacement=False, meta_step_size=0.5, meta_batch_size=mb, lr_ph=m.lr_ph, lr=None)
eng = m.engine()
torch.cuda.synchronize()
th = eng.tf_order_vector(eng.theta(0)).clone()
assert torch.isfinite(th).all()
t = th.clone()
dist.broadcast(t, 0)
assert torch.equal(t, th), "theta diverged between ranks after meta-steps with idle ranks"
bn = eng.bn_state(0).clone()
t = bn.clone()
dist.broadcast(t, 0)
assert torch.equal(t, bn), "BN statistics diverged between ranks"
print("[rank %d] slot-parallel meta-steps with idle ranks: theta and BN statistics identical across ranks" % rank)
dist.barrier()
dist.destroy_process_group()
