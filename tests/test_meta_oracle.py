"""oracle/meta_oracle.py (and the host mirror's Session path) against golden vectors produced by the REFERENCE'S OWN
`Gecko.train_step` / `FOMLIS.train_step` / samplers / variable arithmetic, executed from their source on a toy session
(tests/golden/make_golden_meta.py -> tests/golden/meta_steps_toy.json).  CPU only."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import meta_oracle as MO
from oracle.efficientlab_oracle import OptState
from tests.golden.make_golden_meta import NP, ToySession, ToyTask, ToyVariableState, toy_grad

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "meta_steps_toy.json")))


class ToyOrc:
    """Duck-typed stand-in for EfficientLabOracle: the toy model's analytic gradient + moving statistic."""

    def loss_and_grad(self, theta, bn, x, y, dc_masks=None):
        g, new = toy_grad(theta.numpy(), bn.numpy(), x.numpy(), y.numpy())
        return None, torch.from_numpy(g), torch.from_numpy(new), None


def test_oracle_samplers_match_reference_sequences():
    gold = json.load(open(os.path.join(GOLD, "sampler_sequences.json")))

    class T:
        batch_size, name = 64, "t"
    for case in gold:
        random.seed(case["seed"])
        for rec in case["tasks"]:
            _, rows = MO.sample_task([T()], case["shots"] + case["test_shots"])
            train, test = MO.split_train_test(rows, case["test_shots"])
            batches = [list(b) for b in MO.mini_batches(train, case["batch"], case["iters"], case["replacement"])]
            assert train == rec["train"] and test == rec["test"] and batches == rec["batches"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_meta_steps_reproduce_reference_code(case):
    st = MO.MetaState(torch.linspace(-1.0, 1.0, NP, dtype=torch.float64), torch.zeros(NP, dtype=torch.float64),
                      OptState(NP, torch.float64, sgd=case["sgd"]))
    dataset = [ToyTask(t) for t in range(4)]
    orc = ToyOrc()
    random.seed(123)
    for rec in case["steps"]:
        kw = dict(lr=case["lr"], default_lr=case["default_lr"], weight_decay_rate=0.9 if case["decay"] else None)
        if case["foml"]:
            MO.fomaml_train_step(orc, st, dataset, case["num_shots"], case["inner_batch"], case["inner_iters"],
                                 case["replacement"], case["meta_step_size"], case["meta_batch"],
                                 tail_shots=case["tail_shots"], **kw)
        else:
            MO.reptile_train_step(orc, st, dataset, case["num_shots"], case["inner_batch"], case["inner_iters"],
                                  case["replacement"], case["meta_step_size"], case["meta_batch"], **kw)
        np.testing.assert_allclose(st.theta.numpy(), rec["theta"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(st.bn.numpy(), rec["stat"], rtol=1e-13, atol=1e-15)
        if not case["sgd"]:
            np.testing.assert_allclose(st.opt.v.numpy(), rec["v"], rtol=1e-13, atol=1e-300)
            assert abs(st.opt.b2p - rec["b2p"]) < 1e-15
    assert random.random() == case["rng_after"]        # identical consumption of the `random` stream


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_host_session_path_reproduces_reference_code(case):
    """mliis_b200.reptile's Session path (the line-by-line mirror) on the same toy session."""
    from mliis_b200.reptile import FOMLIS, Gecko
    sess = ToySession(case["sgd"], case["default_lr"])
    cls = FOMLIS if case["foml"] else Gecko
    learner = cls.__new__(cls)                       # no engine: only the control flow is under test
    learner.session = sess
    learner._model_state = ToyVariableState(sess)
    learner._pre_step_op = "decay" if case["decay"] else None
    learner.lr_scheduler = None
    learner.augmenter = None
    learner.aug_rate = None
    learner.fast_path = False
    if case["foml"]:
        learner.tail_shots = case["tail_shots"]
        learner.train_shots = case["num_shots"] - case["tail_shots"] if case["tail_shots"] is not None else case["num_shots"]
        learner.sample_train_val_with_replacement = False
    dataset = [ToyTask(t) for t in range(4)]
    random.seed(123)
    for rec in case["steps"]:
        learner.train_step(dataset, "X", "Y", "minimize", 1, case["num_shots"], case["inner_batch"], case["inner_iters"],
                           case["replacement"], case["meta_step_size"], case["meta_batch"], lr_ph="lr_ph", lr=case["lr"])
        np.testing.assert_allclose(sess.theta, rec["theta"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(sess.stat, rec["stat"], rtol=1e-13, atol=1e-15)
        assert sess.runs == rec["runs"]
    assert random.random() == case["rng_after"]


def test_evaluate_task_restores_state_and_counts():
    """evaluate_task leaves the caller's state untouched (reptile.py:258, :293) and scores with Gecko._iou."""
    class Task(ToyTask):
        def arrays(self):
            return self.x, self.y

    class Orc(ToyOrc):
        def predict(self, theta, bn, x):
            # a fake 2-channel "mask": pixel j of image b is foreground iff theta_j * x_bj > 0
            fg = (theta * x > 0).float().view(x.shape[0], 1, NP)
            return torch.stack([1 - fg, fg], -1), None
    t = Task(1, n=10)
    t.y = np.stack([1.0 - (t.x > 0), (t.x > 0).astype(np.float64)], -1).reshape(10, 1, NP, 2)   # labels [n,1,NP,2]
    # labels feed the toy gradient as y: keep a [n, NP] view for it
    yy = t.y

    class Task2(Task):
        def arrays(self_inner):
            return t.x, yy
    st = MO.MetaState(torch.ones(NP, dtype=torch.float64), torch.zeros(NP, dtype=torch.float64),
                      OptState(NP, torch.float64))
    before = st.clone()

    class Orc2(Orc):
        def loss_and_grad(self, theta, bn, x, y, dc_masks=None):
            return super().loss_and_grad(theta, bn, x, y[:, 0, :, 1])
    random.seed(0)
    miou, counts = MO.evaluate_task(Orc2(), st, Task2(1, n=10), 5, 5, 8, 2, False)
    assert torch.equal(st.theta, before.theta) and torch.equal(st.bn, before.bn) and st.opt.b2p == before.opt.b2p
    assert len(counts) == 5 and all(0 <= i <= u <= NP for i, u in counts) and 0.0 <= miou <= 1.0
