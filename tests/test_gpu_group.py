"""Task-batched execution (mliis_task_args.n_group): G task slots adapt their tasks in
lockstep, every kernel launched once with the slot as a grid dimension.  With the single-slot partition of the
reduction partials (MLIIS_GROUP_CANONICAL=1) the result must be BIT-IDENTICAL to single-slot runs; with the default,
group-sized partition (fewer, longer wgrad pixel ranges / row chunks per slot) it is the same sum in a different
order: deterministic for a given group size, equal to rounding."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZE = 64


def _plans(n, T, B, first=2000, with_dc=None):
    from mliis_b200 import metaseg
    from mliis_b200.runner import TaskPlan
    from mliis_b200.synthetic import SyntheticSegmentationTask
    plans = []
    rng = np.random.default_rng(0)
    for t in range(n):
        task = SyntheticSegmentationTask(first + t, 10, SIZE)
        _, rows = metaseg._sample_task_indices([task], 10)
        train, test = metaseg._split_train_test_segmentation(rows, 5)
        batches = list(metaseg._mini_batches(train, B, T, False))
        images, labels = task.arrays()
        dc = (rng.random((T, with_dc, B)) > 0.2).astype(np.float32) if with_dc else None
        plans.append(TaskPlan(images, labels, np.asarray(batches, np.int32), np.full(T, 1e-3, np.float32),
                              np.asarray(test, np.int32), dc, task.name))
    return plans


@pytest.mark.parametrize("with_dc,dropout", [(False, 0.0), (True, 0.5)])
def test_task_groups_are_bit_identical_to_single_slots(with_dc, dropout, monkeypatch):
    from mliis_b200 import native as N
    monkeypatch.setenv("MLIIS_GROUP_CANONICAL", "1")        # read by the launchers at every call (kernels.h partition_nz)
    from mliis_b200.engine import Engine
    from mliis_b200.pretrain import synthetic_checkpoint
    from mliis_b200.runner import TaskRunner
    T, B = 3, 8
    eng = Engine(image_size=SIZE, max_batch=8, n_slots=4, gemm_mode=N.GEMM_TF32X3, final_dropout_rate=dropout)
    init = synthetic_checkpoint(eng, steps=40)
    random.seed(3)
    plans = _plans(8, T, B, with_dc=eng.n_dc if with_dc else None)
    out = {}
    for G in (1, 2, 4):
        runner = TaskRunner(eng, 10, T, B, 5, use_graph=True, with_dc_masks=with_dc, group=G)
        runner.set_init_state(init)
        res = runner.run(plans)
        torch.cuda.synchronize()
        out[G] = (res, eng.states.clone())
        # a short last chunk (6 plans on groups of 4) pads with a repeated plan and still returns every result
        runner._tasks_staged = 0         # same staged dropout seeds as the first run
        res6 = runner.run(plans[:6])
        assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(res6, res[:6]))
    for G in (2, 4):
        for a, b in zip(out[1][0], out[G][0]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert torch.equal(out[1][1], out[G][1]), "adapted slot states differ between group sizes 1 and %d" % G
    inter = np.stack([r[0] for r in out[1][0]])
    assert inter.sum() > 0            # the checkpoint segments: the comparison is not 0 == 0
    # eager (no graph) group execution gives the same answer
    runner = TaskRunner(eng, 10, T, B, 5, use_graph=False, with_dc_masks=with_dc, group=4)
    runner.set_init_state(init)
    res = runner.run(plans)
    for a, b in zip(out[1][0], res):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_group_sized_partition_is_deterministic_and_equal_to_rounding(monkeypatch):
    """Default partition of the reduction partials (sized by the group): the adapted states of group sizes 1 and 4 agree
    to fp32 rounding of the weight-gradient sums, the IoU counts to a handful of threshold pixels, and a repeated run of
    the same group size is bit-identical."""
    from mliis_b200 import native as N
    from mliis_b200.engine import Engine
    from mliis_b200.pretrain import synthetic_checkpoint
    from mliis_b200.runner import TaskRunner
    monkeypatch.delenv("MLIIS_GROUP_CANONICAL", raising=False)
    T, B = 3, 8
    eng = Engine(image_size=SIZE, max_batch=8, n_slots=4, gemm_mode=N.GEMM_TF32X3)
    init = synthetic_checkpoint(eng, steps=40)
    random.seed(3)
    plans = _plans(8, T, B)
    out = {}
    for G in (1, 4, 4):
        runner = TaskRunner(eng, 10, T, B, 5, use_graph=True, group=G)
        runner.set_init_state(init)
        res = runner.run(plans)
        torch.cuda.synchronize()
        if G in out:
            for a, b in zip(out[G][0], res):
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            assert torch.equal(out[G][1], eng.states)
        out[G] = (res, eng.states.clone())
    th1, th4 = out[1][1][:, :eng.n_theta].double(), out[4][1][:, :eng.n_theta].double()
    rel = float((th1 - th4).norm() / th1.norm())
    print("group 1 vs 4, group-sized partition: theta relL2 %.2e" % rel)
    assert rel < 2e-4          # Adam with beta1 = 0 turns rounding of a near-zero gradient into a +-lr step (measured 4.7e-5)
    i1 = np.stack([r[0] for r in out[1][0]]).astype(np.int64)
    i4 = np.stack([r[0] for r in out[4][0]]).astype(np.int64)
    assert i1.sum() > 0
    assert np.abs(i1 - i4).sum() <= max(8, 5e-3 * i1.sum())


def test_task_group_validation():
    from mliis_b200 import native as N
    from mliis_b200.engine import Engine
    from mliis_b200.runner import TaskRunner
    eng = Engine(image_size=SIZE, max_batch=8, n_slots=3, gemm_mode=N.GEMM_TF32X3)
    with pytest.raises(ValueError):
        TaskRunner(eng, 10, 2, 8, 5, group=2)          # 2 does not divide 3 slots
    eng32 = Engine(image_size=SIZE, max_batch=8, n_slots=2, gemm_mode=N.GEMM_FP32)
    with pytest.raises(ValueError):
        TaskRunner(eng32, 10, 2, 8, 5, group=2)        # fp32 FFMA mode serves one slot per launch
