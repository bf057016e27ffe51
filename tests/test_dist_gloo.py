"""world_size-2 run of the task-parallel host logic over the gloo backend (CPU)."""
import os
import random

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mliis_b200 import metaseg
    from mliis_b200.reptile import _dist
    from mliis_b200.runner import allreduce_meta, gather_owned, owned_indices
    from mliis_b200.synthetic import SyntheticSegmentationTask
    assert _dist() == (rank, world)
    # 1. every rank draws every task: identical index plans and identical `random` state on all ranks
    random.seed(0)
    tasks = [SyntheticSegmentationTask(i, 10, 32) for i in range(7)]
    plans = []
    for t in tasks:
        obj, rows = metaseg._sample_task_indices([t], 10)
        train, test = metaseg._split_train_test_segmentation(rows, 5)
        plans.append((obj.name, train, test, [list(b) for b in metaseg._mini_batches(train, 8, 5)]))
    gathered = [None] * world
    dist.all_gather_object(gathered, (plans, random.random()))
    assert all(g == gathered[0] for g in gathered)
    # 2. ownership is a partition; results come back complete and identical on every rank
    mine = owned_indices(len(plans), rank, world)
    vals = np.full(len(plans), np.nan)
    for i in mine:
        vals[i] = 0.1 * i + 0.01          # stand-in for the task IoU computed on this rank's GPU
    full = gather_owned(vals)
    assert np.allclose(full, 0.1 * np.arange(len(plans)) + 0.01)
    # 3. meta-update exchange: sum of per-rank delta sums == single-process sum; BN statistics are averaged
    g = torch.Generator().manual_seed(123)
    deltas = torch.randn(5, 1000, generator=g, dtype=torch.float32)       # 5 tasks of a meta-batch
    dsum = torch.zeros(1000)
    for t in owned_indices(5, rank, world):
        dsum += deltas[t]
    bn = torch.full((2, 16), float(rank + 1))
    allreduce_meta(dsum, bn)
    assert torch.allclose(dsum, deltas.sum(0), atol=1e-5)
    assert torch.allclose(bn, torch.full((2, 16), (1 + world) / 2.0 if world == 2 else 1.0))
    theta = torch.ones(1000) + (0.1 / 5) * dsum
    out = [None] * world
    dist.all_gather_object(out, theta.numpy().tobytes())
    assert all(o == out[0] for o in out)           # replicated theta stays bit-identical across ranks
    with open(os.path.join(out_dir, "ok_%d" % rank), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


def test_two_rank_task_parallel_logic(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok_%d" % r)) for r in range(world))
