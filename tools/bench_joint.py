#!/usr/bin/env python
"""Joint-training throughput (BASELINE config 5: EfficientLab, 1000 classes + background, 224x224, batch 32 per GPU,
SGD, L2, dropout 0.2; data-parallel under torchrun with one gradient all-reduce per step).

    python tools/bench_joint.py [--batch 32] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_joint.py

Prints one JSON line: images/s over all ranks, device timed (CUDA events, max over ranks), inputs uploaded from
host memory inside the timed region (that is the product path: JointTrainer.train_step)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--gemm-mode", default="tf32x3")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from mliis_b200 import joint_train as jt
    from mliis_b200.efficientlab import EfficientLab
    from mliis_b200.synthetic import make_task_arrays
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    n_ex = 64
    iu8, mu8 = [], []
    for t in range(8):
        i, m = make_task_arrays(t, n_ex // 8, a.size)
        iu8.append(i)
        mu8.append(m)
    rng = np.random.default_rng(0)
    data = jt.SparseSegmentationData(np.concatenate(iu8), np.concatenate(mu8),
                                     rng.integers(1, a.classes + 1, n_ex).astype(np.int32), a.classes)
    batcher = jt.SparseBatcher(data, a.batch * world, seed=0, rank=rank, world=world)
    model = EfficientLab(n_classes=a.classes, seperate_background_channel=True, binary_iou_loss=False, n_rows=a.size,
                         n_cols=a.size, rsd=[2, 4], l2=True, final_layer_dropout_rate=0.2, optimizer="sgd",
                         learning_rate=5e-3, gemm_mode=a.gemm_mode, task_slots=1, max_batch=a.batch)
    model.initialize()
    trainer = jt.JointTrainer(model)
    batches = [batcher.next_batch() for _ in range(4)]
    launches0 = 0
    for s in range(a.warmup):
        trainer.train_step(*batches[s % 4], lr=5e-3, seed=s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from mliis_b200 import native as N
    launches0 = N.lib().mliis_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        loss = trainer.train_step(*batches[s % 4], lr=5e-3, seed=100 + s)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = ms.item() / a.steps
        print(json.dumps({"metric": "joint-train images/s (EfficientLab, %d classes + background, %dx%d, SGD)"
                          % (a.classes, a.size, a.size), "value": a.batch * world / (ms_step * 1e-3),
                          "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "dtype": a.gemm_mode,
                          "data": "synthetic", "loss": loss,
                          "gpu_launches": int(N.lib().mliis_launch_count() - launches0),
                          "config": {"workload": "joint_train config 5, batch %d per GPU, inputs from host memory "
                                                 "every step" % a.batch}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
