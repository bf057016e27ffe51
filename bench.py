#!/usr/bin/env python
"""Headline benchmark: meta-test adapted-tasks/s (5-shot, 224x224, EfficientLab-6-3, 5 inner Adam steps).

    python bench.py --gpus N --steps K --warmup W            # B200 engine (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference path (oracle/)

One "step" = one pass of the hot path over one batch of synthetic tasks: every rank adapts + evaluates
`--tasks-per-step` tasks (Gecko._evaluate, reptile.py:235-294: state reset, 5 inner steps of batch 8 on the 5 support
images, transductive prediction of the 5 query images, integer IoU counts).  Tasks shard across ranks with no data-path
collective ("weak" scaling: per-GPU work is fixed).

JSON keys follow the driver contract.  `value` = device-resident throughput (task pools already in HBM); `e2e` = the
same through TaskRunner with HOST task arrays (pinned staging, H2D of every pool, D2H of the counts inside the timed
region); `roofline` = the dominant kernel (conv2d_2 of decode_skip_connections_1) timed alone as the task-batched
launch the graphs issue (`--group` slots per launch; `single_slot` keeps the one-slot launch beside it); `roofline_hbm` =
the HBM-bound kernels timed alone, task-batched, L2 flushed; `tensor_peaks` = the measured bf16 (cuBLAS) and kind::tf32
(own tcgen05 loop) tensor-pipe peaks; `cpu_baseline` = the oracle (a torch-CPU restatement of the reference graph -
TF-1.15 cannot run here) on this box's host cores, on the SAME tasks from the SAME checkpoint, which also gives
`miou_vs_oracle`; `meta_train` = meta-steps/s of FOMAML (meta-batch 5) and Reptile (meta-batch 40) with the ONE
all-reduce of the meta-update through the C ABI (BASELINE configs 3 and 4).  Both arms print the same `config`.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
# one hardware work queue per task group (the default of 8 would alias them onto shared queues); must be set
# before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_SHOTS, N_QUERY, INNER_BATCH, INNER_STEPS, LR, IMAGE_SIZE, POOL = 5, 5, 8, 5, 1e-3, 224, 10
FWD_GFLOP_PER_IMAGE = 3.994937
TASK_GFLOP = 3 * FWD_GFLOP_PER_IMAGE * INNER_BATCH * INNER_STEPS + FWD_GFLOP_PER_IMAGE * N_QUERY   # 499.4
CPU_TASKS = 16                          # "oracle on a 16-task subset"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tasks-per-step", type=int, default=96)
    ap.add_argument("--slots", type=int, default=48)
    ap.add_argument("--group", type=int, default=24, help="task slots per task-batched launch (must divide --slots); measured on\n"
                    "B200 (slots x group): 32 x 8 -> 134.7, 32 x 16 -> 140.1, 48 x 16 -> 139.9, 48 x 24 -> 143.2, 64 x 32 -> 143.3 tasks/s")
    ap.add_argument("--gemm-mode", default="auto", choices=["auto", "fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-meta-train", action="store_true")
    ap.add_argument("--meta-slots", type=int, default=20, help="task slots of the meta-training measurement")
    ap.add_argument("--skip-kernels", action="store_true", help="skip the per-kernel roofline micro-benchmarks")
    ap.add_argument("--profile-region", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region of the device-resident run: `ncu --profile-from-start "
                         "off ...` then lists the launches of the timed steps only (not the checkpoint synthesis / capture)")
    ap.add_argument("--cpu-tasks", type=int, default=CPU_TASKS)
    ap.add_argument("--sgd", action="store_true")
    ap.add_argument("--shots", type=int, default=5, choices=[1, 5],
                    help="support images per task (BASELINE config 2 sweeps 1-shot and 5-shot; the headline is 5-shot)")
    a = ap.parse_args()
    global N_SHOTS
    N_SHOTS = a.shots
    return a


def metric_name():
    return "meta-test adapted-tasks/s (%d-shot, 224x224, EfficientLab-6-3, 5 inner Adam steps)" % N_SHOTS


def workload_config(sgd=False):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": "meta-test sweep on synthetic FSS-1000-shaped tasks: per task full-state reset, %d inner %s steps "
                        "(batch %d drawn from %d support images, lr %g), transductive predict of %d query images, integer "
                        "IoU counts" % (INNER_STEPS, "SGD" if sgd else "Adam", INNER_BATCH, N_SHOTS, LR, N_QUERY),
            "model": "EfficientLab-6-3 (efficientnet-b0 truncated at block 10, rsd 2 4), bce_dice + l2",
            "image_size": IMAGE_SIZE, "shots": N_SHOTS, "query": N_QUERY, "inner_batch": INNER_BATCH,
            "inner_steps": INNER_STEPS, "l2": "inputs larger than L2 (one 10 MB task pool per task, 584 MB workspace per slot)"}


# ----------------------------------------------------------------------------------------------------
def make_plans(n_tasks, first_id):
    """Plans drawn with the reference's own sampler call sequence (metaseg.py) under the caller's random.seed."""
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
# CPU arm: ONE protocol for `cpu_baseline` and `--impl reference`
# ----------------------------------------------------------------------------------------------------
def cpu_oracle_run(plans, state, threads):
    """The CPU restatement of the reference path (oracle/), float32, on the host cores.  plans[0] is the warm-up (one
    inner step, untimed); every other plan is one timed task: full-state reset, 5 inner steps, predict, IoU.
    state: None (reference initialisers) or an oracle MetaState (theta, BN statistics, Adam slots of the checkpoint).
    Returns (seconds per timed task, mIoU per timed task)."""
    import torch
    from oracle.efficientlab_oracle import Arch, EfficientLabOracle, OptState, iou_counts
    from oracle.meta_oracle import MetaState, _minimize
    torch.set_num_threads(threads)
    arch = Arch()
    orc = EfficientLabOracle(arch, torch.float32)
    if state is None:
        state = MetaState(arch.init_theta(0, torch.float32), arch.init_bn_state(torch.float32),
                          OptState(arch.n_params, torch.float32))
    times, mious = [], []
    for i, pl in enumerate(plans):
        t0 = time.perf_counter()
        w = state.clone()                                            # _full_state.export_variables (reptile.py:258)
        x, y = torch.from_numpy(pl.images), torch.from_numpy(pl.labels)
        for s in range(1 if i == 0 else INNER_STEPS):
            idx = torch.from_numpy(pl.batch_index[s].astype(np.int64))
            _minimize(orc, w, x[idx], y[idx], float(pl.lrs[s]))
        q = torch.from_numpy(pl.query_index.astype(np.int64))
        pred, _ = orc.predict(w.theta, w.bn, x[q])
        counts = [iou_counts(pred[j].numpy(), pl.labels[pl.query_index[j]]) for j in range(len(q))]
        if i > 0:
            times.append(time.perf_counter() - t0)
            mious.append(float(np.mean([(a + 1e-7) / (b + 1e-7) for a, b in counts])))
    return times, mious


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle.efficientlab_oracle import Arch  # noqa: F401  (fail early if the oracle is missing)
    random.seed(0)
    plans = make_plans(1 + args.warmup + args.steps, 0)               # plans[0]: one-step warm-up of the oracle itself
    times, _ = cpu_oracle_run(plans, None, threads)
    times = times[args.warmup:]
    total = float(np.sum(times))
    value = len(times) / total
    line = {
        "impl": "reference", "metric": metric_name(), "value": value, "unit": "tasks/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.sgd),
        "cpu_baseline": {"value": value, "unit": "tasks/s", "cores": threads, "kind": "port",
                         "sample": "%d timed tasks (one per step) after %d warm-up tasks; torch-CPU float32 restatement "
                                   "of the reference graph (TF-1.15 cannot run here), all host threads; one process "
                                   "regardless of --gpus" % (len(times), args.warmup)},
        "e2e": {"value": value, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# per-kernel micro-benchmarks (CUDA events on the launching stream, L2 flushed between launches)
# ----------------------------------------------------------------------------------------------------
class _Flusher:
    def __init__(self):
        import torch
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB of L2

    def __call__(self):
        self.buf.zero_()


def _time_launch(fn, flush, reps=8, warm=3):
    import torch
    times = []
    for it in range(warm + reps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if it >= warm:
            times.append(e0.elapsed_time(e1))
    return float(np.mean(times))


def time_dominant_kernel(mode, flush, n_group=1):
    """conv2d_2 of decode_skip_connections_1 (efficientlab.py:224): 3x3, 360->112 at 56x56, B=8 - 57 % of the forward
    FLOPs.  Algorithmic work = the reference op: 2*8*56*56*9*360*112 = 18.21 GFLOP per task slot.  As built: the 136
    image-pooling channels are folded into a per-image, per-border-class bias (pool_taps kernel + the conv's prologue) and
    the implicit GEMM runs over the 224 real channels; both launches are inside the timed region.  n_group > 1: the
    task-batched launch the bench's graphs issue (one launch serves n_group slots, each with its own weights)."""
    import torch
    from mliis_b200 import native as N
    B, H, Cin, Cp, Cout = INNER_BATCH, 56, 224, 136, 112
    lib = N.lib()
    st = torch.cuda.current_stream().cuda_stream
    flops = 2.0 * B * H * H * 9 * (Cin + Cp) * Cout * n_group
    if mode == N.GEMM_FP32:
        g = torch.Generator(device="cuda").manual_seed(0)
        xf = torch.randn(B, H, H, Cin + Cp, device="cuda", generator=g)
        w = torch.randn(3, 3, Cin + Cp, Cout, device="cuda", generator=g) * 0.02
        bias = torch.zeros(Cout, device="cuda")
        y = torch.empty(B, H, H, Cout, device="cuda")
        ms = _time_launch(lambda: N.check(lib.mliis_conv3x3_fwd(xf.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B,
                                                                H, H, Cin + Cp, Cout, 1, mode, st)), flush)
        return ms, flops / n_group
    ar = _Arena(n_group, dict(x=B * H * H * Cin, pooled=B * Cp, w=9 * (Cin + Cp) * Cout, wt=2 * 9 * Cin * Cout, bias=Cout,
                              b9=B * 9 * Cout, y=B * H * H * Cout))
    ar.buf[:, ar.off["w"]:ar.off["w"] + 9 * (Cin + Cp) * Cout] *= 0.04
    N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
    try:
        N.check(lib.mliis_tc_prep_weights_sub(ar.p("w"), ar.p("wt"), 9, Cin, Cin + Cp, Cout, 0, mode, st))
        ms = _time_launch(lambda: N.check(lib.mliis_rsd_conv2_fwd(ar.p("x"), Cin, ar.p("pooled"), ar.p("w"), ar.p("wt"),
                                                                  ar.p("bias"), ar.p("b9"), ar.p("y"), B, H, H, Cin, Cp,
                                                                  Cout, mode, st)), flush)
    finally:
        N.check(lib.mliis_kernel_group(1, 0))
    return ms, flops


class _Arena:
    """n slot copies of a set of named fp32 buffers at one uniform stride (what a task-batched launch addresses)."""

    def __init__(self, n, spec):
        import torch
        self.off, o = {}, 0
        for name, numel in spec.items():
            self.off[name] = o
            o += (int(numel) + 63) // 64 * 64
        self.stride = o
        self.buf = torch.randn(n, o, device="cuda") * 0.5

    def p(self, name):
        return self.buf.data_ptr() + 4 * self.off[name]


def hbm_rooflines(n_group, peak_gbs, flush):
    """The HBM-bound kernels of the path, each timed ALONE at a canonical layer shape, task-batched over n_group slots
    (one launch), L2 flushed.  bytes = algorithmic bytes (unique external inputs + outputs, fp32) x n_group."""
    import torch
    from mliis_b200 import native as N
    lib = N.lib()
    st = torch.cuda.current_stream().cuda_stream
    B = INNER_BATCH
    out = []

    def add(name, ref, nbytes, fn, launches):
        ms = _time_launch(fn, flush)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "replaces": ref, "bytes": int(nbytes), "ms": ms, "GBps": gbs, "frac": gbs / peak_gbs,
                    "launches": launches, "slots_per_launch": n_group})

    # depthwise: 3x3 s1 of blocks_2 (56x56x144), 5x5 s1 of blocks_9 (14x14x672), 3x3 s2 of blocks_1 (112x112x96)
    for (k, s, H, Cc) in ((3, 1, 56, 144), (5, 1, 14, 672), (3, 2, 112, 96)):
        Ho = (H + s - 1) // s
        nx, ny = B * H * H * Cc, B * Ho * Ho * Cc
        scratch = int(lib.mliis_kernel_scratch_floats(B, H, H, Cc))
        ar = _Arena(n_group, dict(x=nx, a=Cc, b=Cc, w=k * k * Cc, y=ny, dy=ny, dx=nx, dw=k * k * Cc, s=scratch))
        N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
        add("dw_fwd %dx%d s%d %dx%dx%d" % (k, k, s, H, H, Cc),
            "DepthwiseConv2dNative + BN+swish prologue (efficientnet_model.py:190-196, :266)", 4.0 * n_group * (nx + ny),
            lambda: N.check(lib.mliis_dwconv_fwd(ar.p("x"), ar.p("w"), ar.p("y"), B, H, H, Cc, k, s, ar.p("a"), ar.p("b"), st)), 1)
        add("dw_bwd (weight + data) %dx%d s%d %dx%dx%d" % (k, k, s, H, H, Cc), "DepthwiseConv2dNativeBackpropFilter/Input",
            4.0 * n_group * (nx + ny + ny + nx),
            lambda: N.check(lib.mliis_dwconv_bwd(ar.p("x"), ar.p("a"), ar.p("b"), ar.p("w"), ar.p("dy"), ar.p("dx"), ar.p("dw"),
                                                 ar.p("s"), B, H, H, Cc, k, s, st)), 3)
        N.check(lib.mliis_kernel_group(1, 0))
        del ar
    # train-mode BN statistics, and backward of swish(BN(x)), on expand outputs (blocks_2: M = 8*56*56, C = 144; blocks_9)
    for (M, Cc) in ((B * 56 * 56, 144), (B * 14 * 14, 672)):
        scratch = int(lib.mliis_kernel_scratch_floats(1, 1, M, Cc))
        ar = _Arena(n_group, dict(x=M * Cc, gamma=Cc, beta=Cc, mm=Cc, mv=Cc, stats=4 * Cc, g=M * Cc, dx=M * Cc, dg=Cc, db=Cc,
                                  s=scratch))
        ar.buf[:, ar.off["mv"]:ar.off["mv"] + Cc] = 1.0
        N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
        add("bn_stats + finalize M=%d C=%d" % (M, Cc), "tf.nn.moments + EMA (utils.py:111-134)", 4.0 * n_group * M * Cc,
            lambda: N.check(lib.mliis_bn_stats_fwd(ar.p("x"), ar.p("gamma"), ar.p("beta"), ar.p("mm"), ar.p("mv"), ar.p("stats"),
                                                   ar.p("s"), M, Cc, 0, st)), 2)
        add("bn_swish_bwd (reduce + finalize + apply) M=%d C=%d" % (M, Cc), "BN + swish gradient (utils.py:87-134)",
            4.0 * n_group * 5 * M * Cc,
            lambda: N.check(lib.mliis_bn_swish_bwd(ar.p("x"), ar.p("g"), ar.p("dx"), ar.p("stats"), ar.p("gamma"), ar.p("dg"),
                                                   ar.p("db"), ar.p("s"), M, Cc, st)), 3)
        N.check(lib.mliis_kernel_group(1, 0))
        del ar
    # squeeze-excite forward of blocks_2 (56x56x144 -> gate)
    HW, Cc, Cr = 56 * 56, 144, 6
    scratch = int(lib.mliis_kernel_scratch_floats(B, 56, 56, Cc))
    ar = _Arena(n_group, dict(x=B * HW * Cc, a=Cc, b=Cc, w1=Cc * Cr, b1=Cr, w2=Cr * Cc, b2=Cc, pool=B * Cc, hid=B * Cr,
                              gate=B * Cc, s=scratch))
    N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
    add("se_pool + se_fc 56x56x144", "squeeze-excite (efficientnet_model.py:238-251)", 4.0 * n_group * B * HW * Cc,
        lambda: N.check(lib.mliis_se_fwd(ar.p("x"), ar.p("a"), ar.p("b"), ar.p("w1"), ar.p("b1"), ar.p("w2"), ar.p("b2"),
                                         ar.p("pool"), ar.p("hid"), ar.p("gate"), ar.p("s"), B, HW, Cc, Cr, st)), 2)
    N.check(lib.mliis_kernel_group(1, 0))
    del ar
    # fused loss at 224x224 (reads low-res logits + 2-channel labels, writes p1, re-reads p1 + labels, writes d logits)
    H, h = IMAGE_SIZE, IMAGE_SIZE // 4
    scratch = int(lib.mliis_kernel_scratch_floats(B, H, H, 4))
    ar = _Arena(n_group, dict(z=B * h * h * 2, y=B * H * H * 2, p1=B * H * H, dz=B * H * H * 2, s=scratch, loss=4))
    yv = ar.buf[:, ar.off["y"]:ar.off["y"] + B * H * H * 2]
    yv.copy_((yv > 0).float())
    N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
    add("loss_fwd + finalize + loss_bwd 224x224", "softmax-CE - ln(dice) + IoU sums + gradient (efficientlab.py:294-327)",
        4.0 * n_group * B * H * H * (2 + 1 + 1 + 2 + 2),
        lambda: N.check(lib.mliis_softmax_ce_iou(ar.p("z"), ar.p("y"), ar.p("p1"), ar.p("dz"), ar.p("s"), ar.p("loss"), B, h, h,
                                                 H, H, 1, 0.0, st)), 3)
    N.check(lib.mliis_kernel_group(1, 0))
    del ar
    # the MBConv pointwise convolutions (tcgen05 path; HBM-bound by bytes at these K): expand 24->144 and project
    # 144->24 of blocks_2 (56x56), project 672->112 of blocks_9/10 (14x14).  The project conv reads the pre-BN
    # depthwise output and applies BN + swish + SE gate in its operand loader (nothing activated is ever stored).
    for (HWs, Cin, Cout, fused) in ((56 * 56, 24, 144, False), (56 * 56, 144, 24, True), (14 * 14, 672, 112, True)):
        M = B * HWs
        ar = _Arena(n_group, dict(x=M * Cin, w=Cin * Cout, wt=2 * Cin * Cout, a=Cin, b=Cin, gate=B * Cin, y=M * Cout))
        N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
        N.check(lib.mliis_tc_prep_weights(ar.p("w"), ar.p("wt"), 1, Cin, Cout, 0, N.GEMM_TF32X3, st))
        if fused:
            fn = lambda: N.check(lib.mliis_tc_project_conv(ar.p("x"), ar.p("wt"), ar.p("a"), ar.p("b"), ar.p("gate"), ar.p("y"),
                                                           B, HWs, Cin, Cout, N.GEMM_TF32X3, st))
            name = "tc_conv 1x1 project %d->%d M=%d (BN+swish+gate prologue)" % (Cin, Cout, M)
            ref = "project conv of MBConvBlock (efficientnet_model.py:225-232, :266, :271-273)"
        else:
            fn = lambda: N.check(lib.mliis_tc_conv(ar.p("x"), ar.p("wt"), None, ar.p("y"), B, 1, HWs, Cin, Cout, 1, 1,
                                                   N.GEMM_TF32X3, st))
            name = "tc_conv 1x1 expand %d->%d M=%d" % (Cin, Cout, M)
            ref = "expand conv of MBConvBlock (efficientnet_model.py:183-188)"
        add(name, ref, 4.0 * n_group * (M * Cin + M * Cout), fn, 1)
        N.check(lib.mliis_kernel_group(1, 0))
        del ar
    # multi-tensor Adam over the flat parameter buffer (reads g, theta, v; writes theta, v)
    P = 2071724
    ar = _Arena(n_group, dict(theta=P, v=P, g=P))
    ar.buf[:, ar.off["v"]:ar.off["v"] + P].abs_()
    N.check(lib.mliis_kernel_group(n_group, ar.stride * 4))
    add("adam_kernel P=2071724", "169 ApplyAdam ops (args.py:151-154)", 4.0 * n_group * 5 * P,
        lambda: N.check(lib.mliis_adam_step(ar.p("theta"), ar.p("v"), ar.p("g"), P, P, 1e-3, 0.999, 0.0005, st)), 2)
    N.check(lib.mliis_kernel_group(1, 0))
    return out


def tensor_peaks(peaks):
    import torch
    from mliis_b200 import native as N
    v = C.c_double()
    N.check(N.lib().mliis_tc_peak_tf32(2048, C.byref(v), torch.cuda.current_stream().cuda_stream))
    return {"bf16_cublas_burst_tflops": peaks["bf16_tflops"], "bf16_cublas_sustained_tflops": peaks.get("bf16_tflops_sustained"),
            "tf32_tcgen05_cta_group1_tflops": v.value,
            "note": "tf32: own tcgen05.mma kind::tf32 M=128 N=256 K=8 loop, operands resident in shared memory, one CTA per "
                    "SM, cta_group::1 (the instruction the convolutions issue); 3xTF32 spends three of these per "
                    "algorithmic MAC"}


# ----------------------------------------------------------------------------------------------------
def meta_train_bench(eng, world, rank, steps, warmup, slots):
    """BASELINE configs 3 and 4 through the host API (FOMLIS / Gecko.train_step on the device fast path): tasks of a
    meta-batch dealt round-robin to ranks and, inside a rank, to `slots` task slots; ONE ncclAllReduce of
    [sum of deltas | BN statistics | count] per meta-step inside the C ABI (mliis_allreduce_delta)."""
    import torch
    import torch.distributed as dist
    from mliis_b200.efficientlab import EfficientLab
    from mliis_b200.reptile import FOMLIS, Gecko
    from mliis_b200.session import Session
    from mliis_b200.synthetic import SyntheticSegmentationTask
    m = EfficientLab(rsd=[2, 4], l2=True, dice=True, final_layer_dropout_rate=0.0, n_rows=IMAGE_SIZE, n_cols=IMAGE_SIZE,
                     learning_rate=LR, optimizer="adam", task_slots=eng.n_slots, gemm_mode="tf32x3")
    m._engine = eng                       # the engine of the meta-test measurement (same configuration)
    m.variables_initialized = True
    sess = Session(m)
    tasks = [SyntheticSegmentationTask(10000 + i, 15, IMAGE_SIZE) for i in range(48)]
    for t in tasks:
        t.arrays()
    out = {}
    for algo, M in (("fomaml", 5), ("reptile", 40)):
        random.seed(0)
        if algo == "fomaml":
            learner = FOMLIS(sess, train_shots=10, tail_shots=5, meta_task_slots=slots)
            kw = dict(num_classes=1, num_shots=10, inner_batch_size=8, inner_iters=5, replacement=False,
                      meta_step_size=0.1, meta_batch_size=M, lr_ph=m.lr_ph, lr=None)
        else:
            learner = Gecko(sess, meta_task_slots=slots)
            kw = dict(num_classes=1, num_shots=5, inner_batch_size=8, inner_iters=5, replacement=False,
                      meta_step_size=0.1, meta_batch_size=M, lr_ph=m.lr_ph, lr=None)
        for _ in range(warmup):
            learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, **kw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # the exchange step alone: 10 all-reduces of the meta buffer back to back
        buf = eng.meta_buffer()
        eng.init_comm()
        for _ in range(3):
            eng.allreduce_delta(buf)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            eng.allreduce_delta(buf)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 10
        if world > 1:
            t = torch.tensor([ar_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ar_ms = float(t.item())
        per = ms / steps
        out[algo] = {"meta_batch": M, "meta_steps_per_s": 1e3 / per, "tasks_per_s": M * 1e3 / per, "ms_per_meta_step": per,
                     "allreduce_ms": ar_ms if world > 1 else 0.0, "allreduce_bytes": int(buf.numel() * 4),
                     "ranks_without_tasks": max(0, world - M), "task_slots_per_rank": slots, "steps": steps,
                     "scaling": "strong (the meta-batch is fixed, its tasks are dealt round-robin to ranks)",
                     "meta_step_gflop": M * 3 * FWD_GFLOP_PER_IMAGE * (37 if algo == "fomaml" else 40)}
    return out


# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from mliis_b200 import native as N
    from mliis_b200.engine import Engine
    from mliis_b200.pretrain import synthetic_checkpoint
    from mliis_b200.runner import TaskPlan, TaskRunner, iou_from_counts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # default: 3xTF32 on the tcgen05 path - the fastest mode that meets every north-star tolerance literally
    mode = {"auto": N.GEMM_TF32X3, "fp32": N.GEMM_FP32, "tf32": N.GEMM_TF32, "tf32x3": N.GEMM_TF32X3}[args.gemm_mode]
    if mode == N.GEMM_FP32:
        args.group = 1                                    # task-batched launches exist on the tensor-core paths only
    if args.slots % args.group:
        raise SystemExit("--group must divide --slots")
    eng = Engine(image_size=IMAGE_SIZE, max_batch=INNER_BATCH, n_slots=args.slots, sgd=args.sgd, gemm_mode=mode,
                 device=local)
    # "checkpoint" (the real tarball is absent): reference initialisers + 100 Adam steps over 8 synthetic "meta-train"
    # tasks + BN recalibration, on the engine (mliis_b200/pretrain.py).  Same on every rank; keeps its optimizer slots.
    random.seed(0)
    init_state = synthetic_checkpoint(eng, steps=100, lr=LR)

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
    if args.profile_region:            # one more step under the profiler (its numbers are not the bench's)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        runner.run(dev_plans)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
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

    meta = None
    if not args.skip_meta_train and mode == N.GEMM_TF32X3 and not args.sgd:
        meta = meta_train_bench(eng, world, rank, steps=max(2, min(args.steps, 6)), warmup=2, slots=min(args.meta_slots, args.slots))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    flush = _Flusher()
    n_grp = max(1, args.group) if mode != N.GEMM_FP32 else 1
    k_ms, k_flops = time_dominant_kernel(mode, flush, n_grp)          # the launch as the bench's graphs issue it
    k1_ms, k1_flops = time_dominant_kernel(mode, flush, 1) if n_grp > 1 else (k_ms, k_flops)
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    tpk = tensor_peaks(peaks) if mode != N.GEMM_FP32 else None
    # tensor-pipe work actually ISSUED per slot: 224 two-row pixel tiles (128 MMA rows for 112 pixels) x 252 k-steps of 8
    # channels x (one 128x224x8 + one 128x112x8 MMA in 3xTF32, one 128x112x8 in TF32)
    issued = 2.0 * 224 * 252 * 128 * 8 * (336 if mode == N.GEMM_TF32X3 else 112) * n_grp
    roof = {"bound": "tensor", "kernel": "conv2d_2 3x3 360->112 @56x56 B=8 (implicit GEMM M=25088 N=112 K=3240 per task slot; as built: "
                                         "pool_taps + tc_conv3_kernel over the 224 real channels, the 136 pooled channels folded into border-class biases), "
                                         "task-batched launch over %d slots as the bench's graphs issue it" % n_grp,
            "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
            "slots_per_launch": n_grp,
            # dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full` capture of this very call at 24 slots per
            # launch (profiles/r02zr_prof_conv3_24slots.raw.csv.gz: tc_conv3_kernel 585.5 MB read + 245.3 MB written,
            # pool_taps_kernel 13.3 MB read = 35.17 MB per slot); algorithmic: 22.5 MB in + 1.8 MB weights + 11.2 MB out
            # = 35.5 MB per slot - nothing is read twice
            "traffic": 35.17e6 * n_grp, "traffic_unit": "bytes per launch (24-slot ncu capture, per slot x slots)",
            "peak_source": "%s bf16 cuBLAS burst (kernel timed alone)" % peak_src,
            "kernel_ms": k_ms, "algorithmic_gflop": k_flops / 1e9,
            "frac_of_measured_tf32_peak": (achieved / tpk["tf32_tcgen05_cta_group1_tflops"]) if tpk else None,
            "tensor_pipe_issued_tflops": (issued / (k_ms * 1e-3) / 1e12) if mode != N.GEMM_FP32 else None,
            "tensor_pipe_issued_frac_of_tf32_peak": (issued / (k_ms * 1e-3) / 1e12 / tpk["tf32_tcgen05_cta_group1_tflops"]) if tpk else None,
            "single_slot": {"kernel_ms": k1_ms, "achieved": k1_flops / (k1_ms * 1e-3) / 1e12,
                            "frac": k1_flops / (k1_ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
                            "note": "one slot per launch: 112 CTAs on 148 SMs"},
            "numeric_mode": {N.GEMM_FP32: "fp32 FFMA", N.GEMM_TF32: "tcgen05 tf32", N.GEMM_TF32X3: "tcgen05 3xtf32"}[mode]}
    hbm = None
    if not args.skip_kernels and mode != N.GEMM_FP32:
        hbm = hbm_rooflines(max(args.group, 6), peaks["hbm_gbs"], flush)
    del flush
    cpu, parity = None, None
    if not args.skip_cpu_baseline and world == 1:       # the CPU arm is timed at N = 1 only (rank 0 would hold the job up)
        threads = os.cpu_count() or 1
        from oracle.meta_oracle import state_from_flat
        eng.states[0].copy_(init_state)
        torch.cuda.synchronize()
        pw = eng.powers(0).cpu()
        st = state_from_flat(eng.tf_order_vector(eng.theta(0)), eng.bn_state(0), eng.tf_order_vector(eng.adam_v(0)),
                             float(pw[0]), float(pw[1]), sgd=args.sgd, dtype=torch.float32)
        n_cpu = min(args.cpu_tasks, tps)
        random.seed(0)
        cpu_plans = make_plans(n_cpu, 1)                           # == host_plans[:n_cpu] of rank 0
        warm = make_plans(1, 0)
        times, or_miou = cpu_oracle_run(warm + cpu_plans, st, threads)
        v = len(times) / float(np.sum(times))
        cpu = {"value": v, "unit": "tasks/s", "cores": threads, "kind": "port",
               "sample": "%d of this run's tasks after a one-step warm-up, from the same checkpoint; torch-CPU float32 "
                         "restatement of the reference graph (oracle/), all host threads; not TF-1.15" % len(times)}
        diff = [abs(a - b) for a, b in zip(mious[:n_cpu], or_miou)]
        parity = {"tasks": n_cpu, "max_abs_diff": float(np.max(diff)), "mean_abs_diff": float(np.mean(diff)),
                  "engine_mean": float(np.mean(mious[:n_cpu])), "oracle_mean": float(np.mean(or_miou)),
                  "oracle_min": float(np.min(or_miou)), "bound": 0.005,
                  "note": "per-task mIoU of the engine (3xTF32) vs the float32 CPU oracle on the same plans from the same state"}
    line = {
        "metric": metric_name(), "value": value, "unit": "tasks/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {N.GEMM_FP32: "f32", N.GEMM_TF32: "tf32", N.GEMM_TF32X3: "tf32x3 (fp32-class)"}[mode], "data": "synthetic",
        "config": workload_config(args.sgd),
        "engine": {"tasks_per_step_per_gpu": tps, "slots": args.slots, "group": args.group, "cuda_graph": not args.no_graph,
                   "parallelism": "task-parallel x%d, no data-path collective" % world,
                   "workspace_mb_per_slot": eng.ctx.workspace_bytes / 2 ** 20},
        "task_gflop": TASK_GFLOP, "achieved_tflops_whole_job": value * TASK_GFLOP / 1e3,
        "mean_iou_check": float(np.mean(mious)), "miou_vs_oracle": parity,
        "roofline": roof, "roofline_hbm": hbm, "tensor_peaks": tpk, "cpu_baseline": cpu, "e2e": e2e, "meta_train": meta,
        "gpu_launches": int(launches), "clocks": sampler.summary(),
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
