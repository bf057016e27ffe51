"""The evaluation metric against the reference's OWN code: tests/golden/iou_metric.json holds the outputs of
`Gecko._iou`, `measure` and `iou_img` executed from the reference's source text (tests/golden/make_golden_iou.py) on seeded
inputs.  Pinned here: the host mirror (mliis_b200.reptile), the oracle's integer counts, and the rule the device kernel
counts with (`predict_kernel`, csrc/k_misc.cu: prediction p1 > 0.5, label > 0.5 == np.round on [0, 1] data, 0.5 -> 0)
followed by `runner.iou_from_counts`."""
import json
import os

import numpy as np

from oracle import efficientlab_oracle as O
from tests.golden.make_golden_iou import cases

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "iou_metric.json")))


def test_host_metric_functions_equal_the_reference():
    from mliis_b200.reptile import Gecko, iou_img, measure
    for name, pred, label in cases():
        g = GOLD[name]
        assert Gecko._iou(pred, label) == g["iou"], name
        assert Gecko._iou(pred, label, round_labels=False) == g["iou_no_label_rounding"], name
        assert Gecko._iou(pred, label, class_of_interest_channel=None) == g["iou_all_channels"], name
        tp, tn, fp, fn = measure(label[:, :, 1], pred[:, :, 1])
        assert [int(tp), int(tn), int(fp), int(fn)] == g["measure"], name
        assert iou_img(tp, fp, fn) == g["iou_img"], name


def test_oracle_counts_and_device_counting_rule_equal_the_reference():
    from mliis_b200.runner import iou_from_counts
    for name, pred, label in cases():
        g = GOLD[name]
        i, u = O.iou_counts(pred, label)
        assert O.iou_score(pred, label) == g["iou"], name
        # what predict_kernel counts: both operands thresholded at 0.5, strictly
        p1, l1 = pred[:, :, 1] > 0.5, label[:, :, 1] > 0.5
        di, du = int(np.logical_and(p1, l1).sum()), int(np.logical_or(p1, l1).sum())
        assert (di, du) == (i, u), name
        assert iou_from_counts(np.array([di], np.uint32), np.array([du], np.uint32)) == g["iou"], name


def test_lr_schedulers_and_normalisation_constants_equal_the_reference():
    """tests/golden/lr_and_constants.json: outputs of the reference's models/lr_schedulers.py and its MEAN_RGB /
    STDDEV_RGB, imported directly (tests/golden/make_golden_sched.py)."""
    import re
    from mliis_b200 import lr_schedulers as L
    from tests.golden.make_golden_sched import COSINE, STEP
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lr_and_constants.json")))
    for (lr, total), ref in zip(COSINE, gold["cosine"]):
        s = L.CosineLRScheduler(lr, total)
        assert [float(s.cur_lr(i)) for i in range(total + 1)] == ref
    for (lr, rate, every), ref in zip(STEP, gold["step"]):
        s = L.StepDecay(lr, None, rate, every)
        assert [float(s.cur_lr(i)) for i in range(12)] == ref
    assert {k: (v.__name__ if v is not None else None) for k, v in L.supported_learning_rate_schedulers.items()} == gold["supported"]
    assert list(O.MEAN_RGB) == gold["MEAN_RGB"] and list(O.STDDEV_RGB) == gold["STDDEV_RGB"]
    # the CUDA constants (csrc/common.cuh) are the float32 roundings of the same products
    src = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "mliis_b200", "csrc", "common.cuh")).read()
    vals = {m.group(1): np.float32(m.group(2)) * np.float32(m.group(3))
            for m in re.finditer(r"(k(?:Mean|Std)[RGB]) = ([0-9.]+)f \* ([0-9.]+)f", src)}
    want = dict(zip(["kMeanR", "kMeanG", "kMeanB", "kStdR", "kStdG", "kStdB"], gold["MEAN_RGB"] + gold["STDDEV_RGB"]))
    assert set(vals) == set(want)
    for k, v in want.items():
        assert abs(float(vals[k]) - v) <= 1e-5 * v, k
