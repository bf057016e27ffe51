"""Task-parallel execution of Gecko._evaluate (reptile.py:235-294) on one GPU.

Every in-flight task owns a slot (state + workspace + staging buffers + CUDA stream).  A task is:
H2D of its example pool and index lists from pinned memory -> ONE CUDA-graph launch (state reset, T inner
steps, transductive prediction, integer IoU counts) -> D2H of the counts.  Slots run concurrently, which is
what fills the 148 SMs: the 14x14 layers of a single task cannot.

The reference does the same work with, per task, 2 full-state host round trips, T+1 feed_dict copies and
~2800 TF op launches from one Python thread.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import native as N
from .engine import Engine, _ptr

import ctypes as C


@dataclass
class TaskPlan:
    """Host-side description of one adapt+evaluate task (all integer work is decided on the host, with the
    reference's own `random` call sequence, so indices are bit-identical to a 1-GPU reference run)."""
    images: np.ndarray                 # [n_pool,S,S,3] f32 in 0..255   (or a device tensor for resident runs)
    labels: np.ndarray                 # [n_pool,S,S,2] f32
    batch_index: np.ndarray            # [T,B] int32 rows of the pool per inner step
    lrs: np.ndarray                    # [T] f32
    query_index: np.ndarray            # [n_query] int32
    dc_mask: Optional[np.ndarray] = None   # [T,n_dc,B] f32 {0,1}; None = all keep
    name: str = ""


class _SlotBuffers:
    """Per-slot task inputs / outputs, carved out of the slot's staging region of the engine arena so that every slot
    has them at the same offset (a task-batched launch addresses slot k as pointer + k * slot_stride)."""

    def __init__(self, eng: Engine, slot: int, n_pool: int, T: int, B: int, nq: int, with_dc: bool,
                 standalone: bool = False):
        S = eng.image_size
        # one small block: [dropout seed (one int64) | batch_index T*B | query nq] int32, [lr T | dc T*n_dc*B] f32
        self.n_i = 2 + T * B + nq
        self.n_f = T + (T * eng.n_dc * B if with_dc else 0)
        stg = eng.staging(slot)
        off = 0

        def carve(n_elems, dtype):
            nonlocal off
            nbytes = n_elems * 4
            if off + nbytes > stg.numel():
                raise ValueError("the engine's per-slot staging region (%d bytes) is too small for this task shape; "
                                 "create the Engine with a larger staging_bytes" % stg.numel())
            t = stg[off:off + nbytes].view(dtype)
            off = (off + nbytes + 255) // 256 * 256
            return t
        pool_bytes = n_pool * S * S * 5 * 4
        if standalone and pool_bytes + (1 << 16) > stg.numel():
            # a pool larger than the arena's staging region (augmented tasks: every inner batch brings its own images);
            # only single-slot launches may keep it outside the uniform-stride arena
            self.images = torch.empty(n_pool, S, S, 3, dtype=torch.float32, device=eng.device)
            self.labels = torch.empty(n_pool, S, S, 2, dtype=torch.float32, device=eng.device)
        else:
            self.images = carve(n_pool * S * S * 3, torch.float32).view(n_pool, S, S, 3)
            self.labels = carve(n_pool * S * S * 2, torch.float32).view(n_pool, S, S, 2)
        self.ints = carve(self.n_i, torch.int32)
        self.floats = carve(max(self.n_f, 1), torch.float32)[:self.n_f]
        self.counts = carve(2 * nq, torch.int32)
        self.losses = carve(T, torch.float32)
        # two pinned host sets per slot: a staging worker fills one while the H2D copies of the other are in flight
        self.host = [_HostSet(S, self.n_i, self.n_f), _HostSet(S, self.n_i, self.n_f)]
        self.h_counts = torch.zeros(2 * nq, dtype=torch.int32).pin_memory()


class _HostSet:
    """Pinned host staging of one task's inputs (example pool grown to the largest pool seen)."""

    def __init__(self, S: int, n_i: int, n_f: int):
        self._S = S
        self.images = self.labels = None
        self.ints = torch.zeros(n_i, dtype=torch.int32).pin_memory()
        self.floats = torch.zeros(n_f, dtype=torch.float32).pin_memory()
        self.h2d_done: Optional[torch.cuda.Event] = None      # recorded after the set's H2D copies were queued

    def ensure(self, n: int) -> None:
        if self.images is None or self.images.shape[0] < n:
            S = self._S
            self.images = torch.empty(n, S, S, 3, dtype=torch.float32).pin_memory()
            self.labels = torch.empty(n, S, S, 2, dtype=torch.float32).pin_memory()


class _Group:
    """`size` consecutive slots that adapt their tasks in lockstep: one stream, one CUDA graph."""

    def __init__(self, eng: Engine, first: int, size: int):
        self.first, self.size = first, size
        self.stream = torch.cuda.Stream(device=eng.device)
        self.done = torch.cuda.Event()
        self.busy = False
        self.tags: List[Optional[int]] = []


class TaskRunner:
    def __init__(self, eng: Engine, n_pool: int = 10, n_steps: int = 5, batch: int = 8, n_query: int = 5,
                 use_graph: bool = True, with_dc_masks: bool = False, pre_decay_rate: float = 1.0, group: int = 1):
        """group: slots per task-batched launch (mliis_task_args.n_group).  group = 1: every slot replays its own
        one-task graph (round-1 behaviour).  group = G > 1: G consecutive slots run in lockstep - each kernel is
        launched once for G tasks - and n_slots / G groups are in flight concurrently.  Results equal single-slot runs to fp32 rounding
        (bit-identical with MLIIS_GROUP_CANONICAL=1: csrc/kernels.h partition_nz)."""
        if batch > eng.max_batch or n_query > eng.max_batch:
            raise ValueError("batch / n_query exceed the engine's max_batch")
        if group < 1 or eng.n_slots % group:
            raise ValueError("group (%d) must divide the engine's n_slots (%d)" % (group, eng.n_slots))
        if group > 1 and eng.gemm_mode == N.GEMM_FP32:
            raise ValueError("task-batched execution needs a tensor-core gemm_mode")
        self.eng = eng
        self.n_pool, self.T, self.B, self.nq = n_pool, n_steps, batch, n_query
        self.use_graph = use_graph
        self.with_dc = with_dc_masks
        self.pre_decay_rate = pre_decay_rate
        self.group = group
        self.slots = [_SlotBuffers(eng, s, n_pool, n_steps, batch, n_query, with_dc_masks, standalone=group == 1)
                      for s in range(eng.n_slots)]
        self.groups = [_Group(eng, g * group, group) for g in range(eng.n_slots // group)]
        self.init_state = torch.zeros(eng.state_floats, dtype=torch.float32, device=eng.device)
        self._captured = False
        self.seed_base = 0          # final-layer dropout: task k of this runner draws its masks from seed_base + k
        self._tasks_staged = 0
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=3, thread_name_prefix="mliis-stage")
        self.h2d_bytes_per_task = 0
        self.d2h_bytes_per_task = 0

    # the shared starting point of every task (= the restored checkpoint, _full_state of reptile.py:258)
    def set_init_state(self, state: torch.Tensor) -> None:
        self.init_state.copy_(state)

    def _task_args(self, g: _Group) -> N.TaskArgs:
        sb = self.slots[g.first]
        T, B, nq = self.T, self.B, self.nq
        dc = sb.floats[T:] if self.with_dc else None
        # the dropout seed is a staged device scalar: every graph replay draws fresh final-layer dropout masks
        return N.TaskArgs(_ptr(self.init_state), _ptr(sb.images), _ptr(sb.labels), _ptr(sb.ints[2:2 + T * B]),
                          _ptr(sb.floats[:T]), T, B, _ptr(sb.ints[2 + T * B:]), nq, _ptr(dc), 0,
                          float(self.pre_decay_rate), _ptr(sb.counts[:nq]), _ptr(sb.counts[nq:]), _ptr(sb.losses),
                          _ptr(sb.ints[:2]), g.size, self.eng.slot_stride if g.size > 1 else 0)

    def _capture(self) -> None:
        lib, h = self.eng.lib, self.eng.ctx.handle
        for sb in self.slots:
            # harmless defaults so that the warm-up run and the capture read valid indices
            sb.ints.zero_()
            if sb.n_f:
                sb.floats.zero_()
            sb.images.zero_()
            sb.labels.zero_()
        torch.cuda.synchronize()
        for gi, g in enumerate(self.groups):
            a = self._task_args(g)
            st = C.c_void_p(g.stream.cuda_stream)
            if gi == 0:   # one eager warm-up sets function attributes before any capture
                N.check(lib.mliis_adapt_eval_task(h, g.first, C.byref(a), st))
                g.stream.synchronize()
            N.check(lib.mliis_task_graph_capture(h, g.first, C.byref(a), st))
        torch.cuda.synchronize()
        self._captured = True

    def _launch(self, g: _Group) -> None:
        lib, h = self.eng.lib, self.eng.ctx.handle
        st = C.c_void_p(g.stream.cuda_stream)
        if self.use_graph:
            N.check(lib.mliis_task_graph_launch(h, g.first, st))
        else:
            a = self._task_args(g)
            N.check(lib.mliis_adapt_eval_task(h, g.first, C.byref(a), st))

    def _fill_host(self, slot: int, which: int, plan: TaskPlan, seed: int) -> None:
        """Host arrays -> the slot's pinned set `which` (runs on a staging worker thread; numpy / torch copies release
        the GIL, so this overlaps the main thread's launches)."""
        torch.cuda.set_device(self.eng.device)       # worker threads start on device 0: pin / wait on the engine's GPU
        sb = self.slots[slot]
        hs = sb.host[which]
        T, B, nq = self.T, self.B, self.nq
        bi = np.asarray(plan.batch_index, np.int32).reshape(-1)
        qi = np.asarray(plan.query_index, np.int32).reshape(-1)
        if bi.size != T * B or qi.size != nq:
            raise ValueError("plan shape mismatch: batch_index %s, query_index %s" % (bi.shape, qi.shape))
        if hs.h2d_done is not None:
            hs.h2d_done.synchronize()          # the copies that last read this set have left the host
        hs.ints[:2].view(torch.int64)[0] = seed
        hs.ints[2:2 + T * B] = torch.from_numpy(bi)
        hs.ints[2 + T * B:] = torch.from_numpy(qi)
        hs.floats[:T] = torch.from_numpy(np.asarray(plan.lrs, np.float32).reshape(-1))
        if self.with_dc:
            dc = plan.dc_mask if plan.dc_mask is not None else np.ones((T, self.eng.n_dc, B), np.float32)
            hs.floats[T:] = torch.from_numpy(np.asarray(dc, np.float32).reshape(-1))
        if not (isinstance(plan.images, torch.Tensor) and plan.images.is_cuda):
            n = plan.images.shape[0]
            if n > self.n_pool:
                raise ValueError("task pool of %d examples exceeds the runner's n_pool (%d)" % (n, self.n_pool))
            hs.ensure(n)
            hs.images[:n] = torch.from_numpy(np.ascontiguousarray(plan.images, np.float32))
            hs.labels[:n] = torch.from_numpy(np.ascontiguousarray(plan.labels, np.float32))

    def _enqueue_h2d(self, slot: int, which: int, plan: TaskPlan, stream) -> None:
        sb = self.slots[slot]
        hs = sb.host[which]
        with torch.cuda.stream(stream):
            h2d = 0
            n = plan.images.shape[0]
            if isinstance(plan.images, torch.Tensor) and plan.images.is_cuda:
                sb.images[:n].copy_(plan.images, non_blocking=True)     # resident pool: device-to-device
                sb.labels[:n].copy_(plan.labels, non_blocking=True)
            else:
                sb.images[:n].copy_(hs.images[:n], non_blocking=True)
                sb.labels[:n].copy_(hs.labels[:n], non_blocking=True)
                h2d += hs.images[:n].numel() * 4 + hs.labels[:n].numel() * 4
            sb.ints.copy_(hs.ints, non_blocking=True)
            if sb.n_f:
                sb.floats.copy_(hs.floats, non_blocking=True)
            h2d += hs.ints.numel() * 4 + hs.floats.numel() * 4
            if hs.h2d_done is None:
                hs.h2d_done = torch.cuda.Event()
            hs.h2d_done.record(stream)
        self.h2d_bytes_per_task = h2d
        self.d2h_bytes_per_task = sb.h_counts.numel() * 4

    def _collect(self, g: _Group, results) -> None:
        g.done.synchronize()
        for k, tag in enumerate(g.tags):
            if tag is None:
                continue
            c = self.slots[g.first + k].h_counts.numpy().copy()
            results[tag] = (c[:self.nq].astype(np.int64), c[self.nq:].astype(np.int64))
        g.busy = False

    def run(self, plans: Sequence[TaskPlan]) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Adapt + evaluate every plan; returns per task (intersection[n_query], union[n_query]) integer counts."""
        if self.use_graph and not self._captured:
            self._capture()
        results: List[Optional[Tuple[np.ndarray, np.ndarray]]] = [None] * len(plans)
        G, ng = self.group, len(self.groups)
        # the group streams are non-blocking: order them after whatever the caller queued on its stream (the copy into
        # init_state, a meta-update still waiting for the training slots) - otherwise a task graph could reset and
        # adapt a slot that a meta-training graph is still running on
        cur = torch.cuda.current_stream()
        for g in self.groups:
            g.stream.wait_stream(cur)
        # host staging runs AHEAD of the launches on worker threads: while chunk ci is being launched, the pinned sets
        # of the next `ng` chunks (one per group, the set the group is not reading from) are already being filled
        chunks = list(range(0, len(plans), G))

        def plan_of(i):     # a short last chunk re-runs the previous plan in the idle slots of the group (discarded)
            return plans[i] if i < len(plans) else plans[len(plans) - 1]

        fills = {}

        def submit_fill(ci):
            g = self.groups[ci % ng]
            which = (ci // ng) & 1      # successive tasks of a group alternate between its two pinned sets
            futs = []
            for k in range(G):
                self._tasks_staged += 1
                futs.append(self._pool.submit(self._fill_host, g.first + k, which, plan_of(chunks[ci] + k),
                                              self.seed_base + self._tasks_staged))
            fills[ci] = (which, futs)

        n_ahead = ng
        for ci in range(min(n_ahead, len(chunks))):
            submit_fill(ci)
        for ci, i0 in enumerate(chunks):
            g = self.groups[ci % ng]
            if g.busy:
                self._collect(g, results)
            if ci + n_ahead < len(chunks):
                submit_fill(ci + n_ahead)
            which, futs = fills.pop(ci)
            g.tags = []
            for k in range(G):
                futs[k].result()                       # re-raises a worker's exception
                i = i0 + k
                self._enqueue_h2d(g.first + k, which, plan_of(i), g.stream)
                g.tags.append(i if i < len(plans) else None)
            self._launch(g)
            with torch.cuda.stream(g.stream):
                for k in range(G):
                    sb = self.slots[g.first + k]
                    sb.h_counts.copy_(sb.counts, non_blocking=True)
                g.done.record(g.stream)
            g.busy = True
        for g in self.groups:
            if g.busy:
                self._collect(g, results)
        return results  # type: ignore


def iou_from_counts(inter: np.ndarray, union: np.ndarray, epsilon: float = 1e-7) -> float:
    """Gecko._iou (reptile.py:549) per image in float64, then np.nanmean over the query set (reptile.py:290-291)."""
    per_image = (inter.astype(np.float64) + epsilon) / (union.astype(np.float64) + epsilon)
    return float(np.nanmean(per_image))


# ---- task-parallel plumbing (one process per GPU; torch.distributed is the transport) ------------------------
def owned_indices(n: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership of n independent units (tasks) by `world` ranks."""
    return [i for i in range(n) if i % world == rank]


def gather_owned(values: np.ndarray, device=None) -> np.ndarray:
    """values: float64 [n] with NaN at indices this rank does not own -> the complete vector on every rank.
    Every index is owned by exactly one rank, so a SUM all-reduce of the NaN-zeroed vectors is exact."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return values
    t = torch.from_numpy(np.nan_to_num(values, nan=0.0))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()


class TrainSlots:
    """S task slots that adapt meta-TRAINING tasks concurrently (Gecko / FOMLIS.train_step, reptile.py:64-125,
    :605-647).  Per slot one CUDA graph: theta_s <- theta_old (trainables only: optimizer slots, beta powers and BN
    moving statistics are NOT reset between tasks, reptile.py:34 vs :102,:123), the inner steps, and
    dsum_s += theta_s - theta_ref (theta_ref = theta_old for Reptile, the weights before the last step for FOMAML).
    Inputs live in per-slot staging buffers so that the graph can be replayed for every task."""

    def __init__(self, eng: Engine, n: int, shape, group: int = 1, group_sizes=None):
        """group = G > 1: G consecutive slots adapt their tasks in LOCKSTEP - one graph per group whose inner steps are
        task-batched launches (mliis_kernel_group + mliis_train_step: every kernel serves G tasks).  The per-slot inputs
        then live in the slots' staging regions of the engine arena (uniform stride).  Equal to group = 1 to fp32 rounding.
        group_sizes = [G0, G1, ..] (sum == n): groups of DIFFERENT sizes on consecutive slots, for task counts that no
        uniform group divides (a meta-batch of 5 as 3 + 2); run_tasks() then takes exactly one task per slot."""
        n_pool, batch_sizes, lrs, fomaml, pre_decay = shape
        if group_sizes is not None:
            group_sizes = [int(g) for g in group_sizes]
            if min(group_sizes) < 1 or sum(group_sizes) != n:
                raise ValueError("group_sizes %r must be positive and sum to the number of training slots (%d)" % (group_sizes, n))
            group = max(group_sizes)
        elif group < 1 or n % group:
            raise ValueError("group (%d) must divide the number of training slots (%d)" % (group, n))
        else:
            group_sizes = [group] * (n // group)
        self.eng, self.n, self.shape, self.fomaml, self.group = eng, n, shape, fomaml, group
        self.uniform = len(set(group_sizes)) == 1
        self.units, s0 = [], 0                     # (first slot, slots in lockstep) of every group
        for g in group_sizes:
            self.units.append((s0, g))
            s0 += g
        self._unit_size = {u[0]: u[1] for u in self.units}
        S, dev = eng.image_size, eng.device
        self.old = torch.empty(eng.n_theta, dtype=torch.float32, device=dev)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
        if group == 1:
            self.x = [torch.empty(n_pool, S, S, 3, dtype=torch.float32, device=dev) for _ in range(n)]
            self.y = [torch.empty(n_pool, S, S, 2, dtype=torch.float32, device=dev) for _ in range(n)]
            self.idx = [[torch.zeros(b, dtype=torch.int32, device=dev) for b in batch_sizes] for _ in range(n)]
            self.seed = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(n)]   # per-replay dropout seeds
        else:
            self.x, self.y, self.idx, self.seed = [], [], [], []
            for s_ in range(n):          # the same carve in every slot: one offset serves the whole group
                stg, off = eng.staging(s_), 0

                def carve(n_elems, dtype):
                    nonlocal off
                    nbytes = n_elems * 4
                    if off + nbytes > stg.numel():
                        raise ValueError("the engine's per-slot staging region is too small for this task shape")
                    t = stg[off:off + nbytes].view(dtype)
                    off = (off + nbytes + 255) // 256 * 256
                    return t
                self.x.append(carve(n_pool * S * S * 3, torch.float32).view(n_pool, S, S, 3))
                self.y.append(carve(n_pool * S * S * 2, torch.float32).view(n_pool, S, S, 2))
                self.idx.append([carve(b, torch.int32) for b in batch_sizes])
                self.seed.append(carve(2, torch.int32).view(torch.int64))
        # two pinned host sets per slot: staging workers fill the set of the slot's NEXT task while the current one runs
        self.xp = [[torch.empty(n_pool, S, S, 3, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(n)]
        self.yp = [[torch.empty(n_pool, S, S, 2, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(n)]
        self.idxp = [[[torch.zeros(b, dtype=torch.int32).pin_memory() for b in batch_sizes] for _ in range(2)]
                     for _ in range(n)]
        self.dsum2d = torch.zeros(n, eng.n_theta, dtype=torch.float32, device=dev)     # per-slot delta sums (rows)
        self.dsum = [self.dsum2d[s] for s in range(n)]
        self.buf = eng.meta_buffer()                      # the exchanged buffer (mliis_meta_reduce / allreduce / finish)
        self.backup = [torch.empty(eng.n_theta, dtype=torch.float32, device=dev) if fomaml else None
                       for _ in range(n)]
        self.seedp = [[torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)] for _ in range(n)]
        self.h2d_done = [[None, None] for _ in range(n)]      # cuda events: the set's copies have left the host
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=3, thread_name_prefix="mliis-train-stage")
        self._submitted = 0
        self.graphs = [None] * n
        self.used = [False] * n
        self._state_ready = False
        self._saved = None           # per-slot training state (optimizer slots, BN statistics) between meta-steps
        self._lrs, self._pre_decay = lrs, pre_decay

    def _body(self, s: int) -> None:
        """One task on each of the slots s .. s + G - 1 (s is the first slot of its group of G)."""
        eng, G = self.eng, self._unit_size[s]
        thetas = [eng.theta(s + k) for k in range(G)]
        for th in thetas:
            th.copy_(self.old)
        T = len(self._lrs)
        for j in range(T):
            if self.fomaml and j == T - 1:
                for k in range(G):
                    self.backup[s + k].copy_(thetas[k])           # last_backup (reptile.py:635-636)
            for k, step_lr in enumerate(self._lrs[j]):
                if G > 1:
                    N.check(eng.lib.mliis_kernel_group(G, eng.slot_stride))
                try:
                    eng.train_step(s, self.x[s], self.y[s], step_lr, index=self.idx[s][j],
                                   pre_decay_rate=self._pre_decay if k == 0 else 1.0, seed=1 + j, seed_dev=self.seed[s])
                finally:
                    if G > 1:
                        N.check(eng.lib.mliis_kernel_group(1, 0))
        for k in range(G):
            eng.delta_accumulate(self.dsum[s + k], thetas[k], self.backup[s + k] if self.fomaml else self.old, first=False)

    def _capture(self, s: int) -> None:
        st = self.streams[s]
        with torch.cuda.stream(st):
            self._body(s)            # warm-up outside the graph (one-time memsets / attribute calls); the slots'
            st.synchronize()         # optimizer / BN state advances by one task: restored by begin() below
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            self._body(s)
        self.graphs[s] = g

    def begin(self, theta_master: torch.Tensor) -> None:
        eng = self.eng
        torch.cuda.synchronize()
        self.old.copy_(theta_master)
        if not self._state_ready:
            # every slot starts from the master's full state (like ranks start from the same checkpoint)
            for s in range(1, self.n):
                eng.states[s].copy_(eng.states[0])
            snap = [eng.states[s].clone() for s in range(self.n)]
            for s in range(self.n):
                self.x[s].zero_()
                self.y[s].zero_()
                for t_ in self.idx[s]:
                    t_.zero_()
                self.seed[s].zero_()
            torch.cuda.synchronize()
            for s, _ in self.units:
                self._capture(s)
            torch.cuda.synchronize()
            for s in range(self.n):
                eng.states[s].copy_(snap[s])
            self._state_ready = True
        elif self._saved is not None:
            # the engine's slots 1.. are shared with the meta-TEST runner (Gecko.evaluate adapts evaluation tasks on
            # them): put back the optimizer slots / BN statistics this meta-trainer left there after its last step
            for s in range(1, self.n):
                eng.states[s].copy_(self._saved[s - 1])
        for s in range(self.n):
            self.dsum[s].zero_()
            self.used[s] = False
        torch.cuda.synchronize()

    def _fill(self, s: int, w: int, images: np.ndarray, labels: np.ndarray, batches, seed: int) -> int:
        """Host arrays -> pinned set w of slot s (runs on a staging worker thread)."""
        torch.cuda.set_device(self.eng.device)
        if self.h2d_done[s][w] is not None:
            self.h2d_done[s][w].synchronize()              # the copies that last read this set have left the host
        n = images.shape[0]                                # <= n_pool (augmented pools vary in size)
        if n > self.xp[s][w].shape[0]:
            raise ValueError("task pool of %d examples exceeds the slot's capacity (%d)" % (n, self.xp[s][w].shape[0]))
        self.xp[s][w][:n].copy_(torch.from_numpy(np.ascontiguousarray(images, np.float32)))
        self.yp[s][w][:n].copy_(torch.from_numpy(np.ascontiguousarray(labels, np.float32)))
        for j, b in enumerate(batches):
            self.idxp[s][w][j].copy_(torch.as_tensor(np.asarray(b, np.int32)))
        self.seedp[s][w][0] = seed
        return n

    def _launch(self, s: int, w: int, counts, n_batches: int) -> None:
        """H2D of the pinned sets of the slots s .. s + group - 1 + ONE graph replay on the group's stream.  No host
        synchronisation: the stream orders the copies after the group's previous tasks (which read the same buffers)."""
        st = self.streams[s]
        with torch.cuda.stream(st):
            for k, n in enumerate(counts):
                q = s + k
                self.seed[q].copy_(self.seedp[q][w], non_blocking=True)
                self.x[q][:n].copy_(self.xp[q][w][:n], non_blocking=True)
                self.y[q][:n].copy_(self.yp[q][w][:n], non_blocking=True)
                for j in range(n_batches):
                    self.idx[q][j].copy_(self.idxp[q][w][j], non_blocking=True)
                if self.h2d_done[q][w] is None:
                    self.h2d_done[q][w] = torch.cuda.Event()
                self.h2d_done[q][w].record(st)
                self.used[q] = True
            self.graphs[s].replay()

    def run_tasks(self, plans) -> None:
        """plans: [(images, labels, batches)] of this rank's tasks of one meta-batch, dealt in chunks of `group` tasks
        round-robin to the groups of slots.  Host staging runs one round AHEAD of the launches (round r uses pinned
        set r & 1 of every slot)."""
        if not self.uniform:             # groups of different sizes: one round, one task per slot
            if len(plans) != self.n:
                raise ValueError("%d tasks for %d slots in groups of %r" % (len(plans), self.n, [u[1] for u in self.units]))
            futs = []
            for s0, g in self.units:
                fl = []
                for k in range(g):
                    images, labels, batches = plans[s0 + k]
                    self._submitted += 1
                    fl.append(self._pool.submit(self._fill, s0 + k, 0, images, labels, batches, 1000003 * self._submitted))
                futs.append(fl)
            for (s0, g), fl in zip(self.units, futs):
                self._launch(s0, 0, [f.result() for f in fl], len(plans[s0][2]))
            return
        G = self.group
        if len(plans) % G:
            raise ValueError("%d tasks do not fill groups of %d slots" % (len(plans), G))
        n_chunks, n_units = len(plans) // G, self.n // G
        # balanced rounds: 10 chunks on 8 units would run as 8 + 2; use ceil(chunks / rounds) units instead (5 + 5)
        rounds = (n_chunks + n_units - 1) // n_units
        nu = (n_chunks + rounds - 1) // rounds if n_chunks else n_units
        futs = {}

        def submit_fill(c):
            s0, w = (c % nu) * G, (c // nu) & 1
            fl = []
            for k in range(G):
                images, labels, batches = plans[c * G + k]
                self._submitted += 1
                fl.append(self._pool.submit(self._fill, s0 + k, w, images, labels, batches, 1000003 * self._submitted))
            futs[c] = fl

        for c in range(min(nu, n_chunks)):
            submit_fill(c)
        for c in range(n_chunks):
            counts = [f.result() for f in futs.pop(c)]
            self._launch((c % nu) * G, (c // nu) & 1, counts, len(plans[c * G][2]))
            if c + nu < n_chunks:
                submit_fill(c + nu)

    def submit(self, s: int, images: np.ndarray, labels: np.ndarray, batches) -> None:
        """One task on slot s, staged synchronously (kept for callers that feed tasks one by one)."""
        if self.group != 1:
            raise ValueError("submit() feeds single slots; use run_tasks() with grouped slots")
        self._submitted += 1
        cnt = self._fill(s, 0, images, labels, batches, 1000003 * self._submitted)
        self._launch(s, 0, [cnt], len(batches))

    def finish(self) -> torch.Tensor:
        """Builds the exchange buffer of this rank on the current stream - [sum of the slot deltas | sum of the BN
        moving statistics of the slots that ran | their count] (mliis_meta_reduce) - and restores the master
        trainables (slot 0) to theta_old.  No host synchronisation, no torch arithmetic."""
        eng = self.eng
        cur = torch.cuda.current_stream()
        n_used = 0
        for s in range(self.n):
            if self.used[s]:
                assert s == n_used, "slots are dealt round-robin from 0"
                cur.wait_stream(self.streams[s])
                n_used += 1
        eng.meta_reduce(self.buf, self.dsum2d, eng.n_theta, 0, n_used)
        eng.theta(0).copy_(self.old)
        return self.buf

    def save_states(self) -> None:
        """After mliis_meta_finish wrote the averaged BN statistics: remember the per-slot training state."""
        eng = self.eng
        if self.n > 1:
            if self._saved is None:
                self._saved = torch.empty(self.n - 1, eng.state_floats, dtype=torch.float32, device=eng.device)
            for s in range(1, self.n):
                self._saved[s - 1].copy_(eng.states[s])


def allreduce_meta(delta_sum: torch.Tensor, bn_state: Optional[torch.Tensor] = None) -> None:
    """HOST-LOGIC reference of the exchange step (used by the CPU / gloo tests only): SUM of the per-rank task deltas
    and the average of the BN moving statistics.  The GPU path does this with ONE ncclAllReduce
    inside the C ABI: mliis_meta_reduce -> mliis_allreduce_delta -> mliis_meta_finish (Engine.allreduce_delta)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    dist.all_reduce(delta_sum)
    if bn_state is not None:
        dist.all_reduce(bn_state)
        bn_state.div_(dist.get_world_size())
