"""The reference-facing Python surface on the GPU: Session idioms, VariableState, Gecko / FOMLIS on both the
Session path and the device fast path, evaluate_gecko / train_gecko, checkpoint save + restore."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZE = 64


def _model(**kw):
    from mliis_b200.efficientlab import EfficientLab
    args = dict(rsd=[2, 4], l2=True, dice=True, final_layer_dropout_rate=0.0, n_rows=SIZE, n_cols=SIZE,
                learning_rate=1e-3, label_smoothing=0.0, optimizer="adam", task_slots=3)
    args.update(kw)
    m = EfficientLab(**args)
    m.initialize(seed=0)
    return m


def _tasks(n, first=0, n_examples=10):
    from mliis_b200.synthetic import SyntheticSegmentationTask
    return [SyntheticSegmentationTask(first + i, n_examples, SIZE) for i in range(n)]


def _warm(sess, m, steps=3):
    t = _tasks(1, 900)[0]
    s = t.sample(sess, 8)
    x, y = zip(*s)
    for _ in range(steps):
        sess.run(m.minimize_op, feed_dict={m.input_ph: x, m.label_ph: y})


def test_session_idioms_and_variable_state():
    from mliis_b200.session import Session
    from mliis_b200.variables import VariableState, weight_decay
    m = _model()
    sess = Session(m)
    vs = VariableState(sess, m.trainable_variables())
    full = VariableState(sess, m.global_variables())
    v0 = vs.export_variables()
    assert len(v0) == 169 and v0[0].shape == (3, 3, 3, 32)
    f0 = full.export_variables()
    t = _tasks(1)[0]
    x, y = zip(*t.sample(sess, 8))
    sess.run(m.minimize_op, feed_dict={m.input_ph: x, m.label_ph: y, m.lr_ph: 1e-3})
    v1 = vs.export_variables()
    assert any(not np.array_equal(a, b) for a, b in zip(v0, v1))
    pred = sess.run(m.predictions, feed_dict={m.input_ph: x[:5], m.is_training_ph: False})
    assert pred.shape == (5, SIZE, SIZE, 2) and set(np.unique(pred)) <= {0.0, 1.0}
    full.import_variables(f0)                       # restores weights, BN statistics and optimizer slots
    f0b = full.export_variables()
    assert all(np.array_equal(a, b) for a, b in zip(f0, f0b))
    # pre_step_op: var <- var * rate, applied before the step
    sess.run(weight_decay(0.5))
    sess.run(m.minimize_op, feed_dict={m.input_ph: x, m.label_ph: y, m.lr_ph: 0.0})
    v2 = vs.export_variables()
    assert np.allclose(v2[0], 0.5 * v0[0], atol=1e-7)
    with pytest.raises(ValueError):
        sess.run(m.minimize_op, feed_dict={m.input_ph: [np.zeros((SIZE, SIZE))], m.label_ph: y})


@pytest.mark.parametrize("shots", [5, 1])
def test_evaluate_fast_path_equals_session_path(shots):
    """5-shot and 1-shot (BASELINE config 2; with one support image every inner batch is that image repeated,
    metaseg.py:288-302)."""
    from mliis_b200.reptile import Gecko
    from mliis_b200.session import Session
    m = _model()
    sess = Session(m)
    _warm(sess, m)
    tasks = _tasks(5, 100)
    kw = dict(num_classes=1, num_shots=shots, inner_batch_size=8, inner_iters=3, replacement=False, eval_all_tasks=True,
              is_training_ph=m.is_training_ph, lr_ph=m.lr_ph)
    before = m.engine().states[0].clone()
    random.seed(11)
    fast = Gecko(sess, transductive=True, fast_path=True)
    mi_f, map_f = fast.evaluate(tasks, m.input_ph, m.label_ph, m.minimize_op, m.predictions, **kw)
    after_rng = random.random()
    assert torch.equal(before, m.engine().states[0])        # evaluation leaves the model state untouched
    random.seed(11)
    slow = Gecko(sess, transductive=True, fast_path=False)
    mi_s, map_s = slow.evaluate(tasks, m.input_ph, m.label_ph, m.minimize_op, m.predictions, **kw)
    assert after_rng == random.random()                     # same consumption of the `random` stream
    assert torch.equal(before, m.engine().states[0])
    assert map_f.keys() == map_s.keys()
    for k in map_f:
        assert abs(map_f[k] - map_s[k]) < 1e-12, (k, map_f[k], map_s[k])
    assert abs(mi_f - mi_s) < 1e-12


@pytest.mark.parametrize("foml", [False, True])
def test_meta_train_step_fast_equals_session_path(foml):
    from mliis_b200.reptile import FOMLIS, Gecko
    from mliis_b200.session import Session
    results = []
    for fast in (True, False):
        m = _model(optimizer="sgd")
        sess = Session(m)
        tasks = _tasks(6, 200)
        random.seed(3)
        if foml:
            learner = FOMLIS(sess, train_shots=10, tail_shots=5, fast_path=fast)
        else:
            learner = Gecko(sess, fast_path=fast)
        for _ in range(2):
            learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, num_classes=1, num_shots=10 if foml else 5,
                               inner_batch_size=4, inner_iters=3, replacement=False, meta_step_size=0.5,
                               meta_batch_size=3, lr_ph=m.lr_ph, lr=None)
        eng = m.engine()
        results.append((eng.tf_order_vector(eng.theta(0)).cpu().double(), eng.bn_state(0).cpu().double(), random.random()))
    (tf_, bf, rf), (ts, bs, rs) = results
    assert rf == rs
    rel = ((tf_ - ts).norm() / ts.norm()).item()
    assert rel < 1e-6, rel                  # host numpy mean/interpolation vs fused device kernels: rounding only
    assert ((bf - bs).abs().max() / bs.abs().max()).item() < 1e-5


@pytest.mark.parametrize("foml,slots", [(False, 1), (True, 1), (True, 3)])
def test_augmented_meta_train_step_fast_equals_session_path(foml, slots):
    """run.sh trains with --augment: the device path stages the augmented copies as pool rows (sequential and
    slot-parallel); same trainables as the Session path under --sgd, same consumption of both RNG streams."""
    from mliis_b200 import np_augmenters
    from mliis_b200.reptile import FOMLIS, Gecko
    from mliis_b200.session import Session
    order = list(np_augmenters.cur_aug_funcs)
    results = []
    for fast in (True, False):
        np_augmenters.cur_aug_funcs[:] = order
        m = _model(optimizer="sgd")
        sess = Session(m)
        tasks = _tasks(6, 250, n_examples=15)
        random.seed(3)
        np.random.seed(3)
        kw = dict(augment=True, aug_rate=0.5, fast_path=fast, meta_task_slots=slots if fast else 1)
        learner = FOMLIS(sess, train_shots=10, tail_shots=5, **kw) if foml else Gecko(sess, **kw)
        for _ in range(2):
            learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, num_classes=1, num_shots=10 if foml else 5,
                               inner_batch_size=4, inner_iters=3, replacement=False, meta_step_size=0.5,
                               meta_batch_size=3, lr_ph=m.lr_ph, lr=None)
        eng = m.engine()
        torch.cuda.synchronize()
        results.append((eng.tf_order_vector(eng.theta(0)).cpu().double(), random.random(), float(np.random.rand())))
    (tf_, rf, nf), (ts, rs, ns) = results
    assert rf == rs and nf == ns
    rel = ((tf_ - ts).norm() / ts.norm()).item()
    assert rel < 1e-5, rel


def test_evaluate_gecko_and_train_gecko_with_checkpoints(tmp_path):
    from mliis_b200.checkpoint import Saver, read_index
    from mliis_b200.eval import evaluate_gecko
    from mliis_b200.session import Session
    from mliis_b200.train import train_gecko
    from mliis_b200.reptile import FOMLIS
    from functools import partial
    from mliis_b200.util import latest_checkpoint
    m = _model()
    sess = Session(m)
    train, test = _tasks(6, 300), _tasks(4, 400)
    random.seed(0)
    save_dir = str(tmp_path / "ckpt")
    meta_fn = partial(FOMLIS, train_shots=10, tail_shots=5)
    train_gecko(sess, m, train, test, save_dir, num_classes=1, num_shots=5, inner_batch_size=4, inner_iters=3,
                meta_step_size=0.1, meta_step_size_final=0.01, meta_batch_size=2, meta_iters=2, eval_inner_batch_size=4,
                eval_inner_iters=2, eval_interval=1, train_shots=10, transductive=True, meta_fn=meta_fn,
                num_tasks_to_eval=2)
    prefix = latest_checkpoint(save_dir)
    assert prefix.endswith("model.ckpt-1")
    idx = read_index(prefix + ".index")
    assert "decode/final_layer_weights/kernel" in idx and "beta2_power" in idx
    assert "efficientnet-b0/model/stem/tpu_batch_normalization/moving_variance" in idx
    mean_iou, iou_map = evaluate_gecko(sess, m, test, num_classes=1, num_shots=5, eval_inner_batch_size=4,
                                       eval_inner_iters=2, num_samples=2, transductive=True, serially_eval_all_tasks=True)
    assert len(iou_map) == 4 and all(len(v) == 2 for v in iou_map.values()) and 0.0 <= mean_iou <= 1.0
    # restore into a fresh model: identical state
    state = m.engine().states[0].clone()
    m2 = _model()
    Saver(m2).restore(Session(m2), prefix)
    s2 = m2.engine().states[0]
    eng = m.engine()
    assert torch.equal(state[:eng.n_theta], s2[:eng.n_theta])
    assert torch.equal(state[eng.o_bn:eng.o_bn + 2 * eng.n_bn], s2[eng.o_bn:eng.o_bn + 2 * eng.n_bn])
    assert torch.equal(state[eng.o_v:eng.o_pow + 2], s2[eng.o_v:eng.o_pow + 2])


def test_early_stopping_k_shot_curves_and_uho(tmp_path):
    """SURVEY 8f-4 drivers on the real engine: early stopping scores the held-out shots after EVERY inner step."""
    from mliis_b200.eval import optimize_update_hyperparams, run_k_shot_learning_curves_experiment
    from mliis_b200.reptile import Gecko
    from mliis_b200.session import Session
    m = _model()
    sess = Session(m)
    _warm(sess, m, 4)
    state0 = m.engine().states[0].clone()
    g = Gecko(sess, transductive=True)
    tasks = _tasks(3, 500, n_examples=12)
    random.seed(1)
    names, steps, ious = g.evaluate_with_early_stopping(
        tasks, m.input_ph, m.label_ph, m.minimize_op, m.predictions, num_classes=1, num_shots=5, inner_batch_size=4,
        min_steps=1, max_steps=6, replacement=False, eval_all_tasks=True, test_shots=5,
        is_training_ph=m.is_training_ph, lr_ph=m.lr_ph, lr=1e-3)
    assert names == [t.name for t in tasks] and len(steps) == 3 and len(ious) == 3
    assert all(1 <= s <= 6 for s in steps) and all(0.0 <= v <= 1.0 for v in ious)
    assert torch.equal(m.engine().states[0], state0)          # every task starts from, and restores, the same state
    # min_steps == max_steps degenerates to a plain evaluation with that many steps
    random.seed(1)
    n2, s2, i2 = g.evaluate_with_early_stopping(
        tasks, m.input_ph, m.label_ph, m.minimize_op, m.predictions, num_classes=1, num_shots=5, inner_batch_size=4,
        min_steps=2, max_steps=2, replacement=False, eval_all_tasks=True, test_shots=5,
        is_training_ph=m.is_training_ph, lr_ph=m.lr_ph, lr=1e-3)
    random.seed(1)
    mean_iou, iou_map = g.evaluate(tasks, m.input_ph, m.label_ph, m.minimize_op, m.predictions, num_classes=1,
                                   num_shots=5, inner_batch_size=4, inner_iters=2, replacement=False,
                                   eval_all_tasks=True, test_shots=5, is_training_ph=m.is_training_ph, lr_ph=m.lr_ph,
                                   lr=1e-3)
    assert s2 == [2, 2, 2] and list(i2) == list(iou_map.values())
    # k-shot learning curves (small k range; 20 % of the shots validate the step count from k = 4 on)
    csv_path = str(tmp_path / "k.csv")
    random.seed(2)
    ks, res = run_k_shot_learning_curves_experiment(
        sess, m, tasks[:1], eval_inner_batch_size=4, eval_inner_iters=2, num_samples=1, lr=1e-3, augment=False,
        csv_outpath=csv_path, k_range=[1, 2, 6], iter_range=[1, 2, 3], test_samples=4)
    assert ks == [1, 2, 6] and len(res) == 3 and all(0.0 <= v <= 1.0 for v in res)
    assert open(csv_path).read().splitlines()[0] == "k,mIoU"
    assert torch.equal(m.engine().states[0], state0)
    # update-hyperparameter optimisation: 3 configurations on 2 validation tasks
    random.seed(3)
    lr, n_steps = optimize_update_hyperparams(
        sess, m, tasks[:2], num_shots=5, eval_inner_batch_size=4, transductive=True, lr=None,
        lr_search_range_low=5e-4, lr_search_range_high=5e-3, drop_rate=None, drop_rate_search_range_low=0.2,
        drop_rate_search_range_high=0.2, min_steps=0, max_steps=3, num_configs_to_sample=3, save_dir=str(tmp_path),
        seed=0)
    assert 5e-4 <= lr <= 5e-3 and 1 <= n_steps <= 3
    assert any(f.startswith("GP_val-set_hyper_param_search_results_5-shot") for f in os.listdir(str(tmp_path)))


def test_augmented_evaluation_on_the_device_fast_path_equals_the_session_path():
    """run.sh evaluates with --augment --aug_rate 0.5 (np_augmenters.py:135-160 through metaseg.py:258-302): the host
    draws the augmentations with the reference's call order on both RNG streams, every augmented copy becomes a row of
    the task's example pool, and the task still runs as one CUDA graph.  Same IoUs and same RNG consumption as the
    Session path; with more tasks than slots the chunked staging is exercised too."""
    from mliis_b200 import np_augmenters
    from mliis_b200.reptile import Gecko
    from mliis_b200.session import Session
    m = _model()
    sess = Session(m)
    _warm(sess, m, 2)
    order = list(np_augmenters.cur_aug_funcs)
    before = m.engine().states[0].clone()
    out = []
    for fast in (True, False):
        np_augmenters.cur_aug_funcs[:] = order      # the reference shuffles this module-level list in place
        g = Gecko(sess, transductive=True, augment=True, aug_rate=0.5, fast_path=fast)
        assert g.augmenter is not None and g.fast_path == fast
        random.seed(0)
        np.random.seed(0)
        mean_iou, iou_map = g.evaluate(_tasks(5, 600), m.input_ph, m.label_ph, m.minimize_op, m.predictions,
                                       num_classes=1, num_shots=5, inner_batch_size=4, inner_iters=3, replacement=False,
                                       eval_all_tasks=True, test_shots=5, is_training_ph=m.is_training_ph,
                                       lr_ph=m.lr_ph, lr=1e-3)
        out.append((mean_iou, iou_map, random.random(), float(np.random.rand())))
        assert torch.equal(before, m.engine().states[0])
    (mf, mapf, rf, nf), (ms, maps, rs, ns) = out
    assert rf == rs and nf == ns
    assert list(mapf.keys()) == list(maps.keys()) and len(mapf) == 5
    for k in mapf:
        assert abs(mapf[k] - maps[k]) < 1e-12, (k, mapf[k], maps[k])
    assert abs(mf - ms) < 1e-12 and 0.0 <= mf <= 1.0


def test_tfrecord_tasks_equal_synthetic_tasks_on_the_device_fast_path(tmp_path):
    """FSS-1000 shard reader: the same records through gzip-TFRecords give bit-identical IoUs."""
    from mliis_b200 import fss1000
    from mliis_b200.reptile import Gecko
    from mliis_b200.session import Session
    from mliis_b200.synthetic import make_task_arrays
    m = _model()
    sess = Session(m)
    _warm(sess, m, 3)
    syn = _tasks(3, 700)
    for t in syn:
        iu8, mu8 = make_task_arrays(t.task_id, 10, SIZE)
        fss1000.write_task_shard(str(tmp_path / (t.name + ".tfrecord.gzip")), iu8, mu8)
    _, _, rec, _, _, names = fss1000.read_fss_1000_dataset(
        str(tmp_path), test_task_ids=[t.name for t in syn], image_size=SIZE)
    rec = sorted(rec, key=lambda t: t.name)
    assert [t.name for t in rec] == [t.name + ".tfrecord.gzip" for t in syn] and all(t.batch_size == 10 for t in rec)
    g = Gecko(sess, transductive=True)
    out = []
    for tasks in (syn, rec):
        random.seed(4)
        out.append(g.evaluate(list(tasks), m.input_ph, m.label_ph, m.minimize_op, m.predictions, num_classes=1,
                              num_shots=5, inner_batch_size=4, inner_iters=3, replacement=False, eval_all_tasks=True,
                              test_shots=5, is_training_ph=m.is_training_ph, lr_ph=m.lr_ph, lr=1e-3))
    assert out[0][0] == out[1][0]
    assert list(out[0][1].values()) == list(out[1][1].values())


@pytest.mark.parametrize("foml,gemm,nslots", [(False, "fp32", 3), (True, "fp32", 3), (False, "tf32x3", 4), (True, "tf32x3", 4),
                                              (True, "tf32x3", 5)])     # 5: two lockstep groups of 3 + 2 slots
def test_slot_parallel_meta_step_equals_sequential_under_sgd(foml, gemm, nslots):
    """meta_task_slots > 1: tasks of a meta-batch adapt concurrently on task slots (one CUDA graph per slot).  With
    SGD the trainables do not depend on the order in which tasks ran (training-mode BN uses batch statistics), so
    the meta-update must equal the sequential reference order up to summation order."""
    from functools import partial
    from mliis_b200.reptile import FOMLIS, Gecko
    from mliis_b200.session import Session
    out, bn = [], []
    # tf32x3 with 4 slots and a meta-batch of 4: the four tasks run as ONE task-batched group (mliis_kernel_group +
    # mliis_train_step, every kernel launched once for the four slots)
    for slots in (1, nslots):
        m = _model(optimizer="sgd", task_slots=nslots, gemm_mode=gemm)
        sess = Session(m)
        _warm(sess, m, 2)
        tasks = _tasks(6, 800, n_examples=15)
        random.seed(11)
        if foml:
            learner = FOMLIS(sess, train_shots=10, tail_shots=5, meta_task_slots=slots)
            kw = dict(num_shots=10, inner_iters=3)
        else:
            learner = Gecko(sess, meta_task_slots=slots)
            kw = dict(num_shots=5, inner_iters=3)
        for _ in range(2):       # two meta-steps: the second replays the captured graphs
            learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, num_classes=1, inner_batch_size=4,
                               replacement=False, meta_step_size=0.5, meta_batch_size=5 if nslots == 5 else 4, lr_ph=m.lr_ph,
                               lr=None, **kw)
        eng = m.engine()
        torch.cuda.synchronize()
        if slots == 5:
            assert [u[1] for u in learner._train_slots.units] == [3, 2]
        out.append(eng.tf_order_vector(eng.theta(0)).double().cpu())
        bn.append(eng.bn_state(0).double().cpu().clone())
        after = random.random()
        out.append(after)
    (th_seq, r_seq, th_par, r_par) = out
    assert r_seq == r_par                                       # identical consumption of the `random` stream
    rel = ((th_par - th_seq).norm() / th_seq.norm()).item()
    assert rel < 1e-5, rel
    assert torch.isfinite(bn[1]).all()


def test_slot_parallel_meta_training_is_not_disturbed_by_evaluation():
    """Gecko.evaluate adapts meta-TEST tasks on the same engine slots the slot-parallel meta-trainer keeps
    its per-slot Adam slots / BN statistics in.  Interleaving evaluations must not change what training computes."""
    from mliis_b200.reptile import Gecko
    from mliis_b200.session import Session
    outs = []
    for with_eval in (False, True):
        m = _model(optimizer="adam", task_slots=3)
        sess = Session(m)
        _warm(sess, m, 2)
        tasks, test_tasks = _tasks(6, 1500, n_examples=10), _tasks(4, 1600)
        learner = Gecko(sess, transductive=True, meta_task_slots=3)
        for k in range(3):
            random.seed(100 + k)
            learner.train_step(tasks, m.input_ph, m.label_ph, m.minimize_op, num_classes=1, num_shots=5,
                               inner_batch_size=4, inner_iters=2, replacement=False, meta_step_size=0.5,
                               meta_batch_size=3, lr_ph=m.lr_ph, lr=None)
            if with_eval:
                learner.evaluate(list(test_tasks), m.input_ph, m.label_ph, m.minimize_op, m.predictions, num_classes=1,
                                 num_shots=5, inner_batch_size=4, inner_iters=2, replacement=False, eval_all_tasks=True,
                                 is_training_ph=m.is_training_ph, lr_ph=m.lr_ph)
        eng = m.engine()
        torch.cuda.synchronize()
        outs.append(eng.states[0].clone())
    assert torch.equal(outs[0], outs[1])


def test_fine_tuned_checkpoints_per_task(tmp_path):
    """--save_fine_tuned_checkpoints (reptile.py:281-285, utils/util.py:72-81): one TF bundle per adapted task."""
    from mliis_b200.checkpoint import read_bundle
    from mliis_b200.eval import evaluate_gecko
    from mliis_b200.session import Session
    m = _model()
    sess = Session(m)
    _warm(sess, m, 2)
    base = m.engine().tf_order_vector(m.engine().theta(0)).cpu().numpy().copy()
    tasks = _tasks(2, 950)
    random.seed(0)
    mean_iou, iou_map = evaluate_gecko(sess, m, tasks, num_shots=5, eval_inner_batch_size=4, eval_inner_iters=2,
                                       num_samples=1, transductive=True, serially_eval_all_tasks=True,
                                       save_fine_tuned_checkpoints=True, save_fine_tuned_checkpoints_dir=str(tmp_path))
    assert len(iou_map) == 2
    for t in tasks:
        prefix = os.path.join(str(tmp_path), t.name, "0", "model.ckpt-1")
        tensors = read_bundle(prefix)
        k = tensors["decode/final_layer_weights/kernel"]
        assert k.shape == (1, 1, 112, 2)
        # the bundle holds the ADAPTED weights of that task, not the meta-learned initialisation ...
        p = [q for q in m.params if q.name == "decode/final_layer_weights/kernel"][0]
        assert not np.allclose(k.reshape(-1), base[_tf_offset(m, p.name):_tf_offset(m, p.name) + k.size])
    # ... and the model itself is back at the initialisation afterwards
    assert np.array_equal(m.engine().tf_order_vector(m.engine().theta(0)).cpu().numpy(), base)


def _tf_offset(m, name):
    off = 0
    for q in m.params:
        if q.name == name:
            return off
        off += q.size
    raise KeyError(name)
