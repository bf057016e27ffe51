"""Golden vectors for the META-STEP control flow, produced by executing the reference's OWN code.

    python tests/golden/make_golden_meta.py        (build container only: needs /root/reference)

The reference's `Gecko.train_step`, `FOMLIS.train_step`, `FOMLIS._mini_batches`, `Gecko._evaluate`-style helpers
cannot run on the real model here (TensorFlow 1.15 is not installable), but their control flow touches TensorFlow
only through `self.session.run(op, feed_dict)` and `VariableState.export/import_variables`.  This script extracts
those methods from /root/reference/meta_learners/supervised_reptile/supervised_reptile/reptile.py by `ast`, the
samplers from meta_learners/metaseg.py and the list arithmetic from meta_learners/variables.py (no TF import is
executed), and runs them against a TOY session: a 6-parameter model with an analytic gradient, an Adam(beta1=0)
optimizer with slots and a BN-like moving statistic that the gradient depends on.  The toy keeps the property that
matters: `_model_state` covers the trainables only, so optimizer slots and the moving statistic flow from task to
task inside a meta-batch (reptile.py:34 vs :102, :123).

Output: tests/golden/meta_steps_toy.json - theta (and the carried state) after every meta-step, for Reptile and
FOMAML (with and without tail shots), Adam and SGD, lr given / not given (the `if / if / else` quirk of
reptile.py:114-121 runs the minimize op twice when lr is given).  tests/test_meta_oracle.py replays the same toy
through oracle/meta_oracle.py and must reproduce these numbers.
"""
import ast
import json
import math
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

NP = 6                      # toy parameters
B1, B2, EPS = 0.0, 0.999, 1e-8


def toy_task_arrays(task_id: int, n: int = 15):
    rng = np.random.default_rng(500 + task_id)
    return rng.standard_normal((n, NP)), rng.standard_normal((n, NP))


def toy_grad(theta, stat, x, y):
    """loss = 0.5 * mean_b |theta * x_b - y_b|^2 + 0.1 * <stat, theta>  ->  (gradient, new moving statistic)."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    g = ((theta * x - y) * x).mean(axis=0) + 0.1 * stat
    new_stat = stat - (stat - x.mean(axis=0)) * (1 - 0.99)
    return g, new_stat


class ToyTask:
    def __init__(self, task_id, n=15):
        self.name = "toy_%d" % task_id
        self.batch_size = n
        self.x, self.y = toy_task_arrays(task_id, n)

    def sample(self, sess, num_images):
        return [[self.x[i], self.y[i]] for i in range(num_images)]

    def arrays(self):
        return self.x, self.y


class ToySession:
    """sess.run(minimize_op, feed_dict) on the toy model: gradient, moving-statistic update, optimizer apply."""

    def __init__(self, sgd: bool, default_lr: float):
        self.theta = np.linspace(-1.0, 1.0, NP)
        self.stat = np.zeros(NP)
        self.v = np.zeros(NP)
        self.b1p, self.b2p = B1, B2
        self.sgd, self.default_lr = sgd, default_lr
        self.runs = 0

    def run(self, op, feed_dict=None):
        if op == "decay":
            self.theta = self.theta * 0.9
            return None
        assert op == "minimize"
        lr = feed_dict.get("lr_ph", self.default_lr)
        g, self.stat = toy_grad(self.theta, self.stat, feed_dict["X"], feed_dict["Y"])
        self.runs += 1
        if self.sgd:
            self.theta = self.theta - lr * g
            return None
        alpha = lr * math.sqrt(1 - self.b2p) / (1 - self.b1p)
        m = B1 * 0.0 + (1 - B1) * g
        self.v = B2 * self.v + (1 - B2) * g * g
        self.theta = self.theta - alpha * m / (np.sqrt(self.v) + EPS)
        self.b1p *= B1
        self.b2p *= B2
        return None


class ToyVariableState:
    """VariableState over the trainables only (meta_learners/variables.py:58-80 contract)."""

    def __init__(self, sess):
        self.sess = sess

    def export_variables(self):
        return [self.sess.theta[:3].copy(), self.sess.theta[3:].copy()]       # a list of arrays, like TF variables

    def import_variables(self, values):
        self.sess.theta = np.concatenate([np.asarray(v, np.float64) for v in values])


def _functions(path, names, class_name=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if class_name is not None:
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0].body
    return [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]


def reference_namespace():
    from typing import Dict, List, Optional, Tuple, Union
    ns = {"random": random, "np": np, "warnings": warnings, "Optional": Optional, "List": List, "Tuple": Tuple,
          "Union": Union, "Dict": Dict, "Augmenter": object, "assert_train_test_split": lambda *a: None}
    mods = [(os.path.join(REF, "meta_learners/metaseg.py"),
             {"_sample_mini_image_segmentation_dataset", "_mini_batches", "_split_train_test_segmentation",
              "_sample_train_test_segmentation_with_replacement"}, None),
            (os.path.join(REF, "meta_learners/variables.py"),
             {"interpolate_vars", "average_vars", "subtract_vars", "add_vars", "scale_vars"}, None)]
    for path, names, cls in mods:
        fns = _functions(path, names, cls)
        assert {f.name for f in fns} == names, (path, names)
        exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    rp = os.path.join(REF, "meta_learners/supervised_reptile/supervised_reptile/reptile.py")
    gecko = _functions(rp, {"train_step"}, "Gecko")
    fomlis = _functions(rp, {"train_step", "_mini_batches"}, "FOMLIS")
    assert len(gecko) == 1 and len(fomlis) == 2
    # methods are compiled INSIDE a class body so that FOMLIS._mini_batches does not shadow metaseg._mini_batches
    def as_class(name, methods):
        cls = ast.ClassDef(name=name, bases=[], keywords=[], body=methods, decorator_list=[])
        mod = ast.Module(body=[cls], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, rp + ":" + name, "exec"), ns)
        return ns[name]
    RefGecko = as_class("RefGecko", gecko)
    RefFOMLIS = as_class("RefFOMLIS", fomlis)
    return RefGecko, RefFOMLIS


CASES = [
    # name, foml, tail_shots, sgd, lr, num_shots, inner_batch, inner_iters, replacement, eps, meta_batch, decay
    ("reptile_adam", False, None, False, None, 5, 8, 3, False, 0.5, 3, False),
    ("reptile_adam_lr_quirk", False, None, False, 2e-3, 5, 8, 3, False, 0.25, 2, False),
    ("reptile_sgd_replacement", False, None, True, None, 10, 4, 4, True, 1.0, 3, False),
    ("reptile_sgd_decay", False, None, True, 5e-3, 5, 8, 2, False, 0.5, 2, True),
    ("fomaml_tail_adam", True, 5, False, None, 10, 8, 4, False, 0.5, 3, False),
    ("fomaml_tail_sgd_lr", True, 5, True, 1e-2, 10, 8, 5, False, 0.1, 5, False),
    ("fomaml_notail_adam", True, None, False, 1e-3, 5, 8, 3, False, 0.5, 2, False),
]
DEFAULT_LR = 1e-3
N_META_STEPS = 3


def run_case(case, RefGecko, RefFOMLIS):
    name, foml, tail, sgd, lr, shots, ib, it, repl, eps, mb, decay = case
    sess = ToySession(sgd, DEFAULT_LR)
    learner = (RefFOMLIS if foml else RefGecko)()
    learner.session = sess
    learner._model_state = ToyVariableState(sess)
    learner._pre_step_op = "decay" if decay else None
    learner.lr_scheduler = None
    learner.augmenter = None
    learner.aug_rate = None
    if foml:
        learner.tail_shots = tail
        learner.train_shots = shots - tail if tail is not None else shots
        learner.sample_train_val_with_replacement = False
    dataset = [ToyTask(t) for t in range(4)]
    random.seed(123)
    steps = []
    for _ in range(N_META_STEPS):
        learner.train_step(dataset, "X", "Y", "minimize", 1, shots, ib, it, repl, eps, mb, lr_ph="lr_ph", lr=lr)
        steps.append({"theta": sess.theta.tolist(), "stat": sess.stat.tolist(), "v": sess.v.tolist(),
                      "b2p": sess.b2p, "runs": sess.runs})
    return {"name": name, "foml": foml, "tail_shots": tail, "sgd": sgd, "lr": lr, "num_shots": shots,
            "inner_batch": ib, "inner_iters": it, "replacement": repl, "meta_step_size": eps, "meta_batch": mb,
            "decay": decay, "default_lr": DEFAULT_LR, "rng_after": random.random(), "steps": steps}


if __name__ == "__main__":
    RefGecko, RefFOMLIS = reference_namespace()
    out = [run_case(c, RefGecko, RefFOMLIS) for c in CASES]
    with open(os.path.join(HERE, "meta_steps_toy.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "meta_steps_toy.json"), [(o["name"], o["steps"][-1]["runs"]) for o in out])
