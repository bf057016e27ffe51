#!/usr/bin/env python
"""Meta-training throughput in meta-steps/s (BASELINE configs 3 and 4) through the host API (FOMLIS / Gecko
.train_step on the device fast path: per-task adaptation on task slots, device-side delta accumulation, one NCCL
all-reduce of sum(delta) + BN statistics per meta-step when world > 1, fused theta += eps/M * sum(delta)).

    python tools/bench_meta.py [--algo fomaml|reptile] [--steps 5] [--warmup 2]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_meta.py ...

fomaml : --foml --foml-tail 5 --train-shots 10 --meta-batch 5  --inner-iters 5 --inner-batch 8   
reptile: --meta-batch 40 --train-shots 5 (= shots) --inner-iters 5 --inner-batch 8                (config 4)
Wall-clock around synchronised steps (the step includes host-side sampling and H2D of every task pool)."""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="fomaml", choices=["fomaml", "reptile"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--meta-batch", type=int, default=None)
    ap.add_argument("--meta-task-slots", type=int, default=8,
                    help="task slots adapting the tasks of a meta-batch concurrently (1 = the reference's sequential order)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from mliis_b200.efficientlab import EfficientLab
    from mliis_b200.reptile import FOMLIS, Gecko
    from mliis_b200.session import Session
    from mliis_b200.synthetic import SyntheticSegmentationTask
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    m = EfficientLab(rsd=[2, 4], l2=True, dice=True, final_layer_dropout_rate=0.0, n_rows=a.size, n_cols=a.size,
                     learning_rate=1e-3, optimizer="adam", task_slots=a.slots, gemm_mode="tf32x3")
    m.initialize(seed=0)
    sess = Session(m)
    tasks = [SyntheticSegmentationTask(10000 + i, 15, a.size) for i in range(64)]
    for t in tasks:
        t.arrays()
    random.seed(0)
    if a.algo == "fomaml":
        M = a.meta_batch or 5
        learner = FOMLIS(sess, train_shots=10, tail_shots=5, meta_task_slots=a.meta_task_slots)
        kw = dict(num_classes=1, num_shots=10, inner_batch_size=8, inner_iters=5, replacement=False,
                  meta_step_size=0.1, meta_batch_size=M, lr_ph=m.lr_ph, lr=None)
    else:
        M = a.meta_batch or 40
        learner = Gecko(sess, meta_task_slots=a.meta_task_slots)
        kw = dict(num_classes=1, num_shots=5, inner_batch_size=8, inner_iters=5, replacement=False,
                  meta_step_size=0.1, meta_batch_size=M, lr_ph=m.lr_ph, lr=None)

    def step():
        learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, **kw)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        per = dt.item() / a.steps
        print(json.dumps({"metric": "meta-steps/s (%s, meta-batch %d, 5 inner Adam steps, batch 8, %dx%d)"
                          % (a.algo, M, a.size, a.size), "value": 1.0 / per, "unit": "meta-steps/s",
                          "tasks_per_s": M / per, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "strong",
                          "dtype": "tf32x3", "data": "synthetic", "meta_task_slots": a.meta_task_slots,
                          "config": {"workload": "%s meta-training step through %s.train_step (host sampling + H2D "
                                                 "inside the timed region), tasks dealt round-robin to ranks"
                                     % (a.algo, "FOMLIS" if a.algo == "fomaml" else "Gecko")}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
