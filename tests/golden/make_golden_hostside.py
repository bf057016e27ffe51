"""Golden vectors for the host-side pieces of SURVEY.md section 8f rows 3-4, produced by RUNNING THE REFERENCE's own
code in the build container (these modules need only numpy / scipy / pandas, not TensorFlow):

    python tests/golden/make_golden_hostside.py        (needs /root/reference; writes tests/golden/hostside_*.{npz,json})

1. hostside_augmenters.npz  - /root/reference/augmenters/np_augmenters.py: every transform and the Augmenter on a seeded
                              24x24 example, with both global RNG streams (np.random, random) seeded.
2. hostside_misc.json       - EarlyStopper decisions (/root/reference/meta_learners/hyperparam_search.py:24-68; skopt is
                              stubbed out, it is only used by the GP search) and split_train_test_tasks
                              (/root/reference/data/fss_1000_utils.py:7-19) under random.seed(0..2).
"""
import importlib
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

AUG_NAMES = ["additive_gaussian_noise", "exposure", "random_eraser", "fliplr", "translate", "rotate_img_mask"]


def example(size=24, seed=0):
    rng = np.random.default_rng(seed)
    image = rng.uniform(0, 255, (size, size, 3)).astype(np.float32)
    fg = (rng.random((size, size)) > 0.6).astype(np.float32)
    mask = np.stack([1 - fg, fg], axis=2).astype(np.float32)
    return image, mask


def seed_all(s):
    np.random.seed(s)
    random.seed(s)


def aug_vectors(mod, out):
    image, mask = example()
    for name in AUG_NAMES:
        for s in range(4):
            seed_all(100 + s)
            im, mk = getattr(mod, name)(image.copy(), mask.copy())
            out["%s_%d_image" % (name, s)] = np.asarray(im)
            out["%s_%d_mask" % (name, s)] = np.asarray(mk)
    # the Augmenter: a run of 8 calls from one seed (the function list is shuffled in place, so calls are coupled)
    by_name = {n: getattr(mod, n) for n in AUG_NAMES}
    order = ["random_eraser", "translate", "fliplr", "additive_gaussian_noise", "exposure", "rotate_img_mask"]
    aug = mod.Augmenter(aug_funcs=[by_name[n] for n in order])
    seed_all(7)
    for i in range(8):
        res = aug.apply_augmentations(image, mask, prob_to_return_original=0.25)
        out["augmenter_%d_image" % i] = np.asarray(res[0])
        out["augmenter_%d_mask" % i] = np.asarray(res[1])
    out["augmenter_final_order"] = np.array([f.__name__ for f in aug.aug_funcs])


def main():
    sys.path.insert(0, REF)
    aug_mod = importlib.import_module("augmenters.np_augmenters")
    out = {}
    aug_vectors(aug_mod, out)
    np.savez_compressed(os.path.join(HERE, "hostside_augmenters.npz"), **out)

    # EarlyStopper: stub skopt (only the GP search uses it)
    skopt = types.ModuleType("skopt")
    skopt.Optimizer = object
    space = types.ModuleType("skopt.space")
    space.Categorical = space.Real = space.Integer = object
    sys.modules["skopt"], sys.modules["skopt.space"] = skopt, space
    hs = importlib.import_module("meta_learners.hyperparam_search")
    misc = {"early_stopper": [], "split": []}
    rng = np.random.default_rng(5)
    for case, (patience, min_steps, increase) in enumerate([(2, 0, True), (3, 4, True), (1, 0, False), (50, 1, True),
                                                            (0, 0, True)]):
        metrics = [float(x) for x in np.round(rng.random(30), 3)]
        es = hs.EarlyStopper(patience, metric_should_increase=increase, min_steps=min_steps)
        decisions = []
        for step, m in enumerate(metrics):
            go = es.continue_training(m, step + 1)
            decisions.append(bool(go))
            if not go:
                break
        misc["early_stopper"].append({"patience": patience, "min_steps": min_steps, "increase": increase,
                                      "metrics": metrics, "decisions": decisions,
                                      "best_metric": es.best_metric(), "best_num_steps": es.best_num_steps()})
    fu = importlib.import_module("data.fss_1000_utils")
    for seed in range(3):
        tasks = ["/d/task_%02d.tfrecord.gzip" % i for i in range(12)]
        random.seed(seed)
        train, test = fu.split_train_test_tasks(list(tasks), 4)
        after = random.random()
        train2, val = fu.split_train_test_tasks(list(train), 2, reproducbile_splits=True)
        misc["split"].append({"seed": seed, "train": train, "test": test, "after": after, "train2": train2,
                              "val": val})
    with open(os.path.join(HERE, "hostside_misc.json"), "w") as f:
        json.dump(misc, f, indent=1)
    print("wrote hostside_augmenters.npz (%d arrays), hostside_misc.json" % len(out))


if __name__ == "__main__":
    main()
