"""Host-side pieces of SURVEY.md 8f rows 3-4 against golden vectors produced by RUNNING the reference's own modules
(tests/golden/make_golden_hostside.py): numpy augmenters, EarlyStopper, the task-split helper.  Bit-exact."""
import json
import os
import random

import numpy as np
import pytest

from mliis_b200 import fss1000, hyperparam_search, np_augmenters

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

AUG_NAMES = ["additive_gaussian_noise", "exposure", "random_eraser", "fliplr", "translate", "rotate_img_mask"]


def _example(size=24, seed=0):
    rng = np.random.default_rng(seed)
    image = rng.uniform(0, 255, (size, size, 3)).astype(np.float32)
    fg = (rng.random((size, size)) > 0.6).astype(np.float32)
    mask = np.stack([1 - fg, fg], axis=2).astype(np.float32)
    return image, mask


def _seed_all(s):
    np.random.seed(s)
    random.seed(s)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "hostside_augmenters.npz"))


@pytest.mark.parametrize("name", AUG_NAMES)
def test_augmentation_matches_reference_bit_exactly(gold, name):
    image, mask = _example()
    for s in range(4):
        _seed_all(100 + s)
        im, mk = getattr(np_augmenters, name)(image.copy(), mask.copy())
        gi, gm = gold["%s_%d_image" % (name, s)], gold["%s_%d_mask" % (name, s)]
        assert np.asarray(im).dtype == gi.dtype and np.asarray(mk).dtype == gm.dtype
        np.testing.assert_array_equal(np.asarray(im), gi)
        np.testing.assert_array_equal(np.asarray(mk), gm)


def test_augmenter_sequence_matches_reference(gold):
    image, mask = _example()
    order = ["random_eraser", "translate", "fliplr", "additive_gaussian_noise", "exposure", "rotate_img_mask"]
    aug = np_augmenters.Augmenter(aug_funcs=[getattr(np_augmenters, n) for n in order])
    _seed_all(7)
    changed = 0
    for i in range(8):
        res = aug.apply_augmentations(image, mask, prob_to_return_original=0.25)
        np.testing.assert_array_equal(np.asarray(res[0]), gold["augmenter_%d_image" % i])
        np.testing.assert_array_equal(np.asarray(res[1]), gold["augmenter_%d_mask" % i])
        changed += int(not np.array_equal(np.asarray(res[0]), image))
    assert changed >= 4                      # the run really exercises the transforms
    assert [f.__name__ for f in aug.aug_funcs] == list(gold["augmenter_final_order"])
    # masks stay one-hot for the label-preserving transforms
    _seed_all(3)
    _, mk = np_augmenters.translate(image.copy(), mask.copy())
    np.testing.assert_array_equal(mk.sum(axis=2), np.ones(mk.shape[:2], np.float32))


def test_early_stopper_matches_reference():
    misc = json.load(open(os.path.join(GOLD, "hostside_misc.json")))
    assert len(misc["early_stopper"]) == 5
    for case in misc["early_stopper"]:
        es = hyperparam_search.EarlyStopper(case["patience"], metric_should_increase=case["increase"],
                                            min_steps=case["min_steps"])
        decisions = []
        for step, m in enumerate(case["metrics"]):
            go = es.continue_training(m, step + 1)
            decisions.append(bool(go))
            if not go:
                break
        assert decisions == case["decisions"]
        assert es.best_metric() == case["best_metric"]
        assert es.best_num_steps() == case["best_num_steps"]


def test_split_train_test_tasks_matches_reference():
    misc = json.load(open(os.path.join(GOLD, "hostside_misc.json")))
    for case in misc["split"]:
        tasks = ["/d/task_%02d.tfrecord.gzip" % i for i in range(12)]
        random.seed(case["seed"])
        train, test = fss1000.split_train_test_tasks(list(tasks), 4)
        assert random.random() == case["after"]
        assert train == case["train"] and test == case["test"]
        train2, val = fss1000.split_train_test_tasks(list(train), 2, reproducbile_splits=True)
        assert train2 == case["train2"] and val == case["val"]


def test_gp_search_finds_the_optimum_of_a_smooth_objective(tmp_path):
    """The skopt stand-in: EI search over (lr log-uniform, batch int) must beat random initialisation."""
    calls = []

    def eval_fn(lr, inner_batch_size, drop_rate, aug_rate, tag):
        calls.append((lr, inner_batch_size))
        score = 1.0 - (np.log10(lr) + 2.0) ** 2 - 0.01 * (inner_batch_size - 6) ** 2     # optimum lr=1e-2, batch 6
        return ["t0", "t1"], [5, 7], [score, score - 0.01]

    params = {"lr": None, "inner_batch_size": 8, "drop_rate": 0.2, "aug_rate": 0.5, "tag": "x"}
    csv_path = str(tmp_path / "search.csv")
    lr, steps = hyperparam_search.lr_droprate_aug_rate_batch_size_gp_search(
        eval_fn, params, lr_search_range_low=0.05, lr_search_range_high=0.0005, batch_size_search_range_low=4,
        batch_size_search_range_high=10, n=16, save_results_to=csv_path, seed=0)
    assert steps == 6
    assert 0.0005 <= lr <= 0.05 and abs(np.log10(lr) + 2.0) < 0.35
    assert all(0.0005 <= c[0] <= 0.05 and 4 <= c[1] <= 10 and isinstance(c[1], int) for c in calls)
    rows = open(csv_path).read().strip().split("\n")
    assert rows[0].split(",")[:3] == ["task_ID", "best_num_steps", "mIoU"] and len(rows) == 1 + 16 * 2
