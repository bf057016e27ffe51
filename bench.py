#!/usr/bin/env python
"""Headline benchmark: meta-test adapted-tasks/s (5-shot, 224x224, EfficientLab-6-3, 5 inner Adam steps).

    python bench.py --gpus N --steps K --warmup W            # B200 engine (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference path (oracle/)

One "step" = one pass of the hot path over one batch of synthetic tasks: every rank adapts + evaluates
`--tasks-per-step` tasks (Gecko._evaluate, reptile.py:235-294: state reset, 5 inner steps of batch 8 on the
5 support images, transductive prediction of the 5 query images, integer IoU counts).  Tasks shard across
ranks with no data-path collective ("weak" scaling: per-GPU work is fixed).

JSON keys follow the driver contract: `value` = device-resident throughput (task pools already in HBM),
`e2e` = the same through TaskRunner with HOST task arrays (pinned staging, H2D of every pool, D2H of the
counts inside the timed region), `roofline` for the dominant kernel, `cpu_baseline` = the oracle (a torch-CPU
restatement of the reference graph - TF-1.15 cannot run here) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per task slot (the default of 8 would alias slots onto shared queues); must be set
# before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_SHOTS, N_QUERY, INNER_BATCH, INNER_STEPS, LR, IMAGE_SIZE, POOL = 5, 5, 8, 5, 1e-3, 224, 10
FWD_GFLOP_PER_IMAGE = 3.994937
TASK_GFLOP = 3 * FWD_GFLOP_PER_IMAGE * INNER_BATCH * INNER_STEPS + FWD_GFLOP_PER_IMAGE * N_QUERY   # 499.4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tasks-per-step", type=int, default=24)
    ap.add_argument("--slots", type=int, default=12)
    ap.add_argument("--group", type=int, default=1, help="task slots per task-batched launch (must divide --slots)")
    ap.add_argument("--gemm-mode", default="auto", choices=["auto", "fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--sgd", action="store_true")
    ap.add_argument("--shots", type=int, default=5, choices=[1, 5],
                    help="support images per task (BASELINE config 2 sweeps 1-shot and 5-shot; the headline is 5-shot)")
    a = ap.parse_args()
    global N_SHOTS
    N_SHOTS = a.shots
    return a


# ----------------------------------------------------------------------------------------------------
def make_plans(n_tasks, first_id, host_arrays=True):
    """Plans drawn with the reference's own sampler call sequence (metaseg.py) under random.seed(0)."""
    from mliis_b200 import metaseg
    from mliis_b200.runner import TaskPlan
    from mliis_b200.synthetic import SyntheticSegmentationTask
    plans = []
    for t in range(n_tasks):
        task = SyntheticSegmentationTask(first_id + t, POOL, IMAGE_SIZE)
        _, rows = metaseg._sample_task_indices([task], N_SHOTS + N_QUERY)
        train, test = metaseg._split_train_test_segmentation(rows, N_QUERY)
        batches = list(metaseg._mini_batches(train, INNER_BATCH, INNER_STEPS, False))
        images, labels = task.arrays()
        plans.append(TaskPlan(images, labels, np.asarray(batches, np.int32), np.full(INNER_STEPS, LR, np.float32),
                              np.asarray(test, np.int32), None, task.name))
    return plans


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu,utilization.memory")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 9:
                    self.samples.append(f)
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons, ug, um, pw = [], [], set(), [], [], []
        for f in self.samples:
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3])); ug.append(float(f[9])); um.append(float(f[10]))
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
               "samples": len(sm)}
        if ug:
            out.update({"util_gpu_pct": float(np.median(ug)), "util_mem_pct": float(np.median(um)),
                        "power_w": float(np.median(pw))})
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------------------------------
def cpu_oracle_tasks_per_s(n_tasks, threads):
    """The CPU restatement of the reference path (oracle/), float32, on the host cores."""
    import torch
    from oracle.efficientlab_oracle import Arch, EfficientLabOracle, OptState, iou_counts
    torch.set_num_threads(threads)
    arch = Arch()
    orc = EfficientLabOracle(arch, torch.float32)
    theta0 = arch.init_theta(0, torch.float32)
    bn0 = arch.init_bn_state(torch.float32)
    plans = make_plans(n_tasks + 1, 0)
    times = []
    for i, pl in enumerate(plans):
        t0 = time.perf_counter()
        th, bn = theta0, bn0
        opt = OptState(arch.n_params, torch.float32)
        x, y = torch.from_numpy(pl.images), torch.from_numpy(pl.labels)
        steps = 1 if i == 0 else INNER_STEPS      # task 0 is the warm-up (one step only)
        for s in range(steps):
            idx = torch.from_numpy(pl.batch_index[s].astype(np.int64))
            _, g, bn, _ = orc.loss_and_grad(th, bn, x[idx], y[idx])
            th = opt.apply(th, g, float(pl.lrs[s]))
        q = torch.from_numpy(pl.query_index.astype(np.int64))
        pred, _ = orc.predict(th, bn, x[q])
        for j in range(len(q)):
            iou_counts(pred[j].numpy(), pl.labels[pl.query_index[j]])
        if i > 0:
            times.append(time.perf_counter() - t0)
    return 1.0 / float(np.median(times)), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.warmup + args.steps
    import torch
    from oracle.efficientlab_oracle import Arch  # noqa: F401  (fail early if the oracle is missing)
    t0 = time.perf_counter()
    tps, times = cpu_oracle_tasks_per_s(n, threads)
    times = times[args.warmup:] if len(times) > args.warmup else times
    total = float(np.sum(times))
    value = len(times) / total
    line = {
        "impl": "reference", "metric": "meta-test adapted-tasks/s (%d-shot, 224x224, EfficientLab-6-3, 5 inner Adam steps)" % N_SHOTS,
        "value": value, "unit": "tasks/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "meta-test 5-shot 224x224: adapt (5 Adam steps, batch 8) + transductive predict of 5 "
                               "query images + IoU, one task per step", "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": value, "unit": "tasks/s", "cores": threads, "kind": "port",
                         "sample": "%d synthetic tasks, torch-CPU float32 restatement of the reference graph "
                                   "(TF-1.15 cannot run here)" % len(times)},
        "e2e": {"value": value, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def time_dominant_kernel(eng, gemm_mode):
    """conv2d_2 of decode_skip_connections_1: 3x3, 360->112 at 56x56, B=8 (57% of the forward FLOPs).
    Timed alone with CUDA events on the launching stream, L2 flushed between launches."""
    import torch
    from mliis_b200 import native as N
    B, H, Cin, Cout = INNER_BATCH, 56, 360, 112
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, H, Cin, device="cuda", generator=g)
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=g) * 0.02
    bias = torch.zeros(Cout, device="cuda")
    y = torch.empty(B, H, H, Cout, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib = N.lib()
    st = torch.cuda.current_stream().cuda_stream
    wt = torch.empty(2 * 9 * Cin * Cout, device="cuda")
    if gemm_mode != N.GEMM_FP32:     # weight re-layout is a separate (tiny) kernel, done once per step in the engine
        N.check(lib.mliis_tc_prep_weights(w.data_ptr(), wt.data_ptr(), 9, Cin, Cout, 0, gemm_mode, st))
    times = []
    for it in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if gemm_mode != N.GEMM_FP32:
            N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout, 9, 1,
                                      gemm_mode, st))
        else:
            N.check(lib.mliis_conv3x3_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout,
                                          1, gemm_mode, st))
        e1.record()
        e1.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    flops = 2.0 * B * H * H * 9 * Cin * Cout
    return ms, flops


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mliis_b200 import native as N
    from mliis_b200.engine import Engine
    from mliis_b200.init import initial_bn_state, initial_variables
    from mliis_b200.runner import TaskPlan, TaskRunner, iou_from_counts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # default: 3xTF32 on the tcgen05 path - the fastest mode that meets every north-star tolerance literally
    mode = {"auto": N.GEMM_TF32X3, "fp32": N.GEMM_FP32, "tf32": N.GEMM_TF32, "tf32x3": N.GEMM_TF32X3}[args.gemm_mode]
    eng = Engine(image_size=IMAGE_SIZE, max_batch=INNER_BATCH, n_slots=args.slots, sgd=args.sgd, gemm_mode=mode,
                 device=local)
    # "checkpoint": random init (reference initialisers) + 100 Adam steps over 8 synthetic "meta-train" tasks, then
    # the BN moving statistics (momentum 0.99: they lag 100s of steps) are replaced by the batch statistics of one
    # image per task so that eval-mode predictions are not degenerate.  Same on every rank.
    random.seed(0)
    eng.init_state(0, initial_variables(eng.ctx.params, 0), *initial_bn_state(eng.n_bn))
    from mliis_b200.synthetic import make_task_arrays, parse_records
    pools = [parse_records(*make_task_arrays(100000 + t, 6, IMAGE_SIZE)) for t in range(8)]    # 8 "meta-train" tasks
    xi = torch.from_numpy(np.concatenate([p[0] for p in pools])).cuda()
    yi = torch.from_numpy(np.concatenate([p[1] for p in pools])).cuda()
    rng = np.random.default_rng(0)
    for s in range(100):
        idx = torch.from_numpy(rng.integers(0, xi.shape[0], INNER_BATCH).astype(np.int32)).cuda()
        eng.train_step(0, xi, yi, LR, index=idx)
    b0 = eng.bn_state(0).clone()
    ridx = torch.arange(0, 48, 6, dtype=torch.int32).cuda()              # one image of each pre-training task
    eng.forward(0, xi, True, index=ridx, want_logits=False)             # one EMA update towards the batch statistics
    torch.cuda.synchronize()
    eng.bn_state(0).copy_(b0 + (eng.bn_state(0) - b0) / (1.0 - 0.99))   # setup-time plumbing, outside any timing
    torch.cuda.synchronize()
    del xi, yi
    init_state = eng.states[0].clone()
    # the checkpoint keeps its optimizer slots: Gecko._full_state covers every global variable (reptile.py:35-36)

    runner = TaskRunner(eng, POOL, INNER_STEPS, INNER_BATCH, N_QUERY, use_graph=not args.no_graph, group=args.group)
    runner.set_init_state(init_state)

    random.seed(0)
    tps = args.tasks_per_step
    host_plans = make_plans(tps, 1 + rank * tps)          # distinct tasks per rank
    dev_plans = [TaskPlan(torch.from_numpy(p.images).cuda(), torch.from_numpy(p.labels).cuda(), p.batch_index, p.lrs,
                          p.query_index, None, p.name) for p in host_plans]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(plans, steps, warmup):
        for _ in range(warmup):
            runner.run(plans)
        barrier()
        launches0 = N.lib().mliis_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for g in runner.groups:
            g.stream.wait_event(e0)
        res = None
        for _ in range(steps):
            res = runner.run(plans)
        for g in runner.groups:
            cur.wait_stream(g.stream)
        e1.record(cur)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, N.lib().mliis_launch_count() - launches0

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms, res, launches = timed(dev_plans, args.steps, max(3, args.warmup))
    sampler.stop()
    value = world * tps * args.steps / (ms / 1e3)
    mious = [iou_from_counts(i, u) for (i, u) in res]

    e2e = None
    if not args.skip_e2e:
        ms_e, res_e, _ = timed(host_plans, args.steps, 1)
        e2e = {"value": world * tps * args.steps / (ms_e / 1e3), "unit": "tasks/s",
               "h2d_bytes_per_step": int(runner.h2d_bytes_per_task * tps),
               "d2h_bytes_per_step": int(runner.d2h_bytes_per_task * tps),
               "api": "mliis_b200.runner.TaskRunner.run(host TaskPlans) -> mliis_task_graph_launch"}
        # same tasks, same plans -> identical integer counts as the resident run
        assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(res, res_e))

    if rank != 0:
        return
    peaks, peak_src = measured_peaks()
    k_ms, k_flops = time_dominant_kernel(eng, mode)
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "conv2d_2 3x3 360->112 @56x56 B=8 (implicit GEMM M=25088 N=112 K=3240)",
            "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r01e_ncu_full_step.md:
            # 39.06 MB read, 0.00 MB written - the 11.2 MB output was still in L2); algorithmic bytes 48.8 MB
            "traffic": 39.06e6 if mode == N.GEMM_TF32X3 else None, "traffic_unit": "bytes per launch",
            "peak_source": "%s bf16 burst (kernel timed alone)" % peak_src,
            "kernel_ms": k_ms, "algorithmic_gflop": k_flops / 1e9,
            "numeric_mode": {N.GEMM_FP32: "fp32 FFMA", N.GEMM_TF32: "tcgen05 tf32", N.GEMM_TF32X3: "tcgen05 3xtf32"}[mode]}
    cpu = None
    if not args.skip_cpu_baseline:
        threads = os.cpu_count() or 1
        v, times = cpu_oracle_tasks_per_s(2, threads)
        cpu = {"value": v, "unit": "tasks/s", "cores": threads, "kind": "port",
               "sample": "2 synthetic tasks after a 1-step warm-up; torch-CPU float32 restatement of the reference "
                         "graph (oracle/), not TF-1.15"}
    line = {
        "metric": "meta-test adapted-tasks/s (%d-shot, 224x224, EfficientLab-6-3, 5 inner Adam steps)" % N_SHOTS,
        "value": value, "unit": "tasks/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {N.GEMM_FP32: "f32", N.GEMM_TF32: "tf32", N.GEMM_TF32X3: "tf32x3 (fp32-class)"}[mode], "data": "synthetic",
        "config": {"workload": "meta-test sweep, 5-shot 224x224 synthetic FSS-1000-shaped tasks: per task state reset, "
                               "5 inner %s steps (batch 8), transductive predict of 5 query images, IoU counts"
                               % ("SGD" if args.sgd else "Adam"),
                   "tasks_per_step_per_gpu": tps, "slots": args.slots, "group": args.group, "cuda_graph": not args.no_graph,
                   "l2": "inputs larger than L2 (%d MB of task pools + %.0f MB workspace per slot)"
                         % (tps * 10, eng.ctx.workspace_bytes / 2 ** 20),
                   "parallelism": "task-parallel x%d, no data-path collective" % world},
        "task_gflop": TASK_GFLOP, "achieved_tflops_whole_job": value * TASK_GFLOP / 1e3,
        "mean_iou_check": float(np.mean(mious)),
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
