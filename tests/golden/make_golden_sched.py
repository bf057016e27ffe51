"""Generates tests/golden/lr_and_constants.json by IMPORTING the reference's own TensorFlow-free modules:

    python tests/golden/make_golden_sched.py        (needs /root/reference)

* models/lr_schedulers.py: `CosineLRScheduler` / `StepDecay` inner-loop learning rates (reptile.py:269-279 asks
  `lr_scheduler.cur_lr(cur_step=step)` for every inner step), for several (initial lr, total steps / decay) settings;
* models/efficientnet/constants.py: MEAN_RGB / STDDEV_RGB of the input normalisation (efficientlab.py:111-119).
"""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


COSINE = [(1e-3, 5), (0.005, 10), (1e-4, 3)]
STEP = [(1e-3, 0.5, 5), (1e-3, 0.1, 1), (0.01, 0.9, 2)]


def main():
    sch = load("models/lr_schedulers.py", "ref_lr_schedulers")
    const = load("models/efficientnet/constants.py", "ref_constants")
    out = {"cosine": [], "step": [], "MEAN_RGB": list(const.MEAN_RGB), "STDDEV_RGB": list(const.STDDEV_RGB),
           "supported": {k: (v.__name__ if v is not None else None) for k, v in sch.supported_learning_rate_schedulers.items()}}
    for lr, total in COSINE:
        s = sch.CosineLRScheduler(lr, total)
        out["cosine"].append([float(s.cur_lr(i)) for i in range(total + 1)])
    for lr, rate, every in STEP:
        s = sch.StepDecay(lr, None, rate, every)
        out["step"].append([float(s.cur_lr(i)) for i in range(12)])
    path = os.path.join(HERE, "lr_and_constants.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
