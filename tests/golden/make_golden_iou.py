"""Generates tests/golden/iou_metric.json by EXECUTING the reference's own metric code (pure numpy) from its source text:

    python tests/golden/make_golden_iou.py        (needs /root/reference)

* `Gecko._iou` (meta_learners/supervised_reptile/supervised_reptile/reptile.py:526-549): the per-image IoU that
  `_evaluate` averages - np.round on predictions and labels (round-half-even: 0.5 -> 0), logical and / or, epsilon;
* `measure` / `iou_img` (reptile.py:555-566): the Shaban et al. cross-check metric.

The inputs are drawn from seeded numpy generators (`cases()` below, imported by the test), only the outputs are stored.
"""
import ast
import json
import os
from typing import Optional, Union

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/meta_learners/supervised_reptile/supervised_reptile/reptile.py"


def cases():
    """[(name, prediction [H,W,2] float32, label [H,W,2] float32)]"""
    out = []
    rng = np.random.default_rng(0)
    for k in range(4):
        h, w = int(rng.integers(5, 40)), int(rng.integers(5, 40))
        fg_p = (rng.random((h, w)) > 0.5).astype(np.float32)
        fg_l = (rng.random((h, w)) > 0.4).astype(np.float32)
        out.append(("binary_%d" % k, np.stack([1 - fg_p, fg_p], 2), np.stack([1 - fg_l, fg_l], 2)))
    z = np.zeros((16, 16), np.float32)
    out.append(("empty_both", np.stack([1 - z, z], 2), np.stack([1 - z, z], 2)))
    o = np.ones((16, 16), np.float32)
    out.append(("full_prediction_empty_label", np.stack([1 - o, o], 2), np.stack([1 - z, z], 2)))
    # soft labels (resized masks): values on a 1/8 grid including exactly 0.5, which np.round sends to 0
    soft = (rng.integers(0, 9, (24, 24)) / 8.0).astype(np.float32)
    fg_p = (rng.random((24, 24)) > 0.5).astype(np.float32)
    out.append(("soft_labels_with_halves", np.stack([1 - fg_p, fg_p], 2), np.stack([1 - soft, soft], 2)))
    return out


def main():
    tree = ast.parse(open(REF).read())
    gecko = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Gecko")
    iou_fn = next(n for n in gecko.body if isinstance(n, ast.FunctionDef) and n.name == "_iou")
    iou_fn.decorator_list = []                                   # @staticmethod: call it as a plain function
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("measure", "iou_img")]
    ns = {"np": np, "Optional": Optional, "Union": Union}
    exec(compile(ast.Module(body=[iou_fn] + fns, type_ignores=[]), "ref_reptile_metrics", "exec"), ns)
    out = {}
    for name, pred, label in cases():
        tp, tn, fp, fn = ns["measure"](label[:, :, 1], pred[:, :, 1])
        out[name] = {"iou": float(ns["_iou"](pred, label)),
                     "iou_no_label_rounding": float(ns["_iou"](pred, label, round_labels=False)),
                     "iou_all_channels": float(ns["_iou"](pred, label, class_of_interest_channel=None)),
                     "measure": [int(tp), int(tn), int(fp), int(fn)], "iou_img": float(ns["iou_img"](tp, fp, fn))}
    path = os.path.join(HERE, "iou_metric.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: round(v["iou"], 4) for k, v in out.items()})


if __name__ == "__main__":
    main()
