"""Generates the committed fixtures under tests/golden/.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

1. args_flags.json   - every flag of the reference CLI with its default / action, extracted textually from
                       /root/reference/meta_learners/args.py (the module itself imports TensorFlow and cannot be
                       imported here).
2. sampler_sequences.json - index sequences produced by the reference's OWN sampler functions
                       (/root/reference/meta_learners/metaseg.py:258-343), executed from their source text with the
                       TensorFlow-dependent imports stripped, under random.seed(0..2).
3. oracle_small.npz  - float64 oracle outputs on a tiny seeded problem (regression pin of the oracle itself;
                       NOT a reference golden: parity is unpinned, see oracle/__init__.py).
"""
import ast
import json
import os
import random
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def args_flags():
    src = open(os.path.join(REF, "meta_learners/args.py")).read()
    tree = ast.parse(src)
    flags = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            name = node.args[0].value
            kw = {}
            for k in node.keywords:
                if k.arg in ("default", "action", "nargs"):
                    try:
                        kw[k.arg] = ast.literal_eval(k.value)
                    except Exception:
                        kw[k.arg] = None
                if k.arg == "type":
                    kw["type"] = getattr(k.value, "id", None)
            flags[name] = kw
    return flags


def sampler_sequences():
    src = open(os.path.join(REF, "meta_learners/metaseg.py")).read()
    tree = ast.parse(src)
    wanted = {"_mini_batches", "_split_train_test_segmentation"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    mod = ast.Module(body=body, type_ignores=[])
    from typing import List, Optional
    ns = {"random": random, "np": np, "Optional": Optional, "List": List, "Augmenter": object,
          "assert_train_test_split": lambda *a: None}
    exec(compile(mod, "ref_metaseg_samplers", "exec"), ns)
    out = []
    for seed in range(3):
        for (shots, test_shots, batch, iters, repl) in [(5, 5, 8, 5, False), (1, 5, 8, 5, False), (10, 5, 8, 59, False),
                                                        (5, 5, 4, 6, True), (10, 5, 8, 4, False)]:
            random.seed(seed)
            rows = list(range(shots + test_shots))
            recs = []
            for _task in range(3):                       # three consecutive tasks share the stream
                random.sample([0], 1)                    # _sample_mini_image_segmentation_dataset's task draw
                train, test = ns["_split_train_test_segmentation"](rows, test_shots)
                if repl and batch > len(train):
                    continue
                batches = [list(b) for b in ns["_mini_batches"](train, batch, iters, repl)]
                recs.append({"train": train, "test": test, "batches": batches})
            out.append({"seed": seed, "shots": shots, "test_shots": test_shots, "batch": batch, "iters": iters,
                        "replacement": repl, "tasks": recs})
    return out


def oracle_small():
    import torch
    from oracle.efficientlab_oracle import EfficientLabOracle, OptState
    from tests.parity_util import make_problem
    arch, theta, bn, images, labels = make_problem(32, 2, task_id=3, theta_seed=1)
    orc = EfficientLabOracle(arch, torch.float64)
    loss, g, nbn, logits = orc.loss_and_grad(theta, bn, torch.from_numpy(images), torch.from_numpy(labels))
    th1 = OptState(arch.n_params, torch.float64).apply(theta, g, 1e-3)
    sel = np.arange(0, arch.n_params, 997)
    return dict(loss=np.float64(loss.item()), logits=logits.numpy()[:, ::4, ::4, :], grad_sel=g.numpy()[sel],
                theta1_sel=th1.numpy()[sel], bn_mean_head=nbn[0, :64].numpy(), bn_var_head=nbn[1, :64].numpy(),
                grad_norm=np.float64(g.norm().item()))


if __name__ == "__main__":
    with open(os.path.join(HERE, "args_flags.json"), "w") as f:
        json.dump(args_flags(), f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "sampler_sequences.json"), "w") as f:
        json.dump(sampler_sequences(), f)
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **oracle_small())
    print("wrote fixtures to", HERE)
