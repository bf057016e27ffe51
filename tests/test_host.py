"""Host-side mirror of the reference surface: CLI flags, samplers (bit-exact index sequences against the
reference's own sampler functions), variable arithmetic, schedulers, checkpoint bundles, synthetic tasks."""
import json
import os
import random

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_cli_flags_match_reference():
    from mliis_b200.args import argument_parser
    gold = json.load(open(os.path.join(GOLD, "args_flags.json")))
    p = argument_parser()
    actions = {a.option_strings[0]: a for a in p._actions if a.option_strings and a.option_strings[0] != "-h"}
    for flag, spec in gold.items():
        assert flag in actions, "missing reference flag %s" % flag
        a = actions[flag]
        if spec.get("action") == "store_true":
            assert a.nargs == 0 and a.default is False, flag
        else:
            assert a.default == spec.get("default"), (flag, a.default, spec.get("default"))
            if spec.get("type"):
                assert a.type.__name__ == spec["type"], flag
            if spec.get("nargs"):
                assert a.nargs == spec["nargs"], flag
    extra = set(actions) - set(gold)
    assert extra == {"--gemm_mode", "--task_slots", "--synthetic_tasks", "--meta_task_slots"}


def test_run_sh_command_line_parses():
    from mliis_b200.args import argument_parser, evaluate_kwargs, model_kwargs, train_kwargs
    cmd = ("--fss_1000 --image_size 224 --pretrained --rsd 2 4 --l2 --foml --foml-tail 5 --final_layer_dropout_rate 0.5 "
           "--augment --aug_rate 0.5 --sgd --loss_name bce_dice --inner-batch 8 --learning-rate 0.0005 --train-shots 10 "
           "--inner-iters 59 --learning_rate_scheduler fixed --meta-iters 50000 --meta-batch 5 --eval-interval 500 "
           "--serially_eval_all_test_tasks --eval-samples 2 --shots 5 --eval-batch 8 --eval-iters 59 --transductive "
           "--model_name efficientlab --sgd --meta-step 0.1 --meta-step-final 0.00001 --checkpoint ck --data-dir d").split()
    a = argument_parser().parse_args(cmd)
    mk, tk, ek = model_kwargs(a), train_kwargs(a), evaluate_kwargs(a)
    assert mk["optimizer"] == "sgd" and mk["rsd"] == [2, 4] and mk["l2"] and "dice" not in mk and mk["n_rows"] == 224
    assert tk["inner_iters"] == 59 and tk["meta_batch_size"] == 5 and tk["train_shots"] == 10
    assert ek["num_samples"] == 2 and ek["transductive"] and ek["lr"] is None
    assert tk["meta_fn"].func.__name__ == "FOMLIS" and tk["meta_fn"].keywords["tail_shots"] == 5
    a2 = argument_parser().parse_args([])
    assert model_kwargs(a2)["dice"] is False and model_kwargs(a2)["optimizer"] == "adam"


def test_samplers_bit_exact_against_reference_functions():
    from mliis_b200 import metaseg
    gold = json.load(open(os.path.join(GOLD, "sampler_sequences.json")))
    assert len(gold) == 15

    class T:
        batch_size = 64
        name = "t"

        def sample(self, sess, n):
            return list(range(n))
    for case in gold:
        random.seed(case["seed"])
        for rec in case["tasks"]:
            _, rows = metaseg._sample_task_indices([T()], case["shots"] + case["test_shots"])
            train, test = metaseg._split_train_test_segmentation(rows, case["test_shots"])
            batches = [list(b) for b in metaseg._mini_batches(train, case["batch"], case["iters"], case["replacement"])]
            assert train == rec["train"] and test == rec["test"] and batches == rec["batches"]
    # the object path consumes the stream identically to the index path
    random.seed(5)
    a = metaseg._sample_mini_image_segmentation_dataset(None, [T()], 1, 7)
    sa = random.random()
    random.seed(5)
    _, b = metaseg._sample_task_indices([T()], 7)
    assert a == b and sa == random.random()


def test_mini_batches_edge_cases():
    from mliis_b200 import metaseg
    with pytest.raises(ValueError):
        list(metaseg._mini_batches([], 8, 2))
    random.seed(0)
    one_shot = list(metaseg._mini_batches([3], 8, 2))          # 1-shot: the single example repeated
    assert one_shot == [[3] * 8, [3] * 8]
    random.seed(0)
    b = list(metaseg._mini_batches(list(range(5)), 8, 5))
    assert all(len(x) == 8 for x in b) and len(b) == 5
    flat = [v for x in b for v in x]
    assert all(sorted(flat[i:i + 5]) == list(range(5)) for i in range(0, 40, 5))   # every epoch is a permutation

    class Small:
        batch_size, name = 3, "s"

        def sample(self, sess, n):
            return list(range(n))
    with pytest.warns(UserWarning):
        assert metaseg._sample_mini_image_segmentation_dataset(None, [Small()], 1, 10) == [0, 1, 2]


def test_variable_arithmetic():
    from mliis_b200.variables import add_vars, average_vars, interpolate_vars, scale_vars, subtract_vars
    rng = np.random.default_rng(0)
    a = [rng.standard_normal((3, 2)), rng.standard_normal(4)]
    b = [rng.standard_normal((3, 2)), rng.standard_normal(4)]
    c = [rng.standard_normal((3, 2)), rng.standard_normal(4)]
    avg = average_vars([a, b, c])
    assert all(np.allclose(x, (p + q + r) / 3) for x, p, q, r in zip(avg, a, b, c))
    it = interpolate_vars(a, b, 0.25)
    assert all(np.allclose(x, p + 0.25 * (q - p)) for x, p, q in zip(it, a, b))
    assert all(np.allclose(x, p + 2 * (q - p)) for x, p, q in zip(add_vars(a, scale_vars(subtract_vars(b, a), 2)), a, b))


def test_lr_schedulers():
    from mliis_b200.lr_schedulers import CosineLRScheduler, StepDecay, supported_learning_rate_schedulers
    c = CosineLRScheduler(1e-3, 10)
    assert abs(c.cur_lr(0) - 1e-3) < 1e-12 and abs(c.cur_lr(10)) < 1e-12 and abs(c.cur_lr(5) - 5e-4) < 1e-12
    s = StepDecay(1e-3, None, 0.5, 5)
    assert s.cur_lr(4) == 1e-3 and s.cur_lr(5) == 5e-4 and s.cur_lr(10) == 2.5e-4 and StepDecay(1e-3, None, 0.1, 1).cur_lr(9) == 1e-7
    assert supported_learning_rate_schedulers["fixed"] is None and set(supported_learning_rate_schedulers) == {
        "cosine_anneal", "fixed", "constant", "step", "step_decay"}


def test_ci95_and_latest_checkpoint(tmp_path):
    from mliis_b200.util import ci95, latest_checkpoint
    a = [0.1, 0.5, 0.9, 0.3]
    assert abs(ci95(a) - 1.96 * np.std(a) / 2) < 1e-12
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-1200"\nall_model_checkpoint_paths: "model.ckpt-1100"\n')
    assert latest_checkpoint(str(tmp_path)) == os.path.join(str(tmp_path), "model.ckpt-1200")


def test_checkpoint_bundle_roundtrip(tmp_path):
    from mliis_b200.checkpoint import crc32c, read_bundle, read_index, write_bundle
    assert crc32c(b"123456789") == 0xE3069283          # CRC-32C check value
    rng = np.random.default_rng(1)
    t = {"efficientnet-b0/model/stem/conv2d/kernel": rng.standard_normal((3, 3, 3, 32)).astype("f4"),
         "decode/final_layer_weights/bias": np.zeros(2, "f4"), "beta2_power": np.float32(0.999)}
    for i in range(150):
        t["efficientnet-b0/model/blocks_%d/x%d/gamma" % (i % 11, i)] = rng.standard_normal(i % 7 + 1).astype("f4")
    prefix = str(tmp_path / "model.ckpt-7")
    write_bundle(prefix, t, with_data_crc=True)
    idx = read_index(prefix + ".index")
    assert idx[""]["num_shards"] == 1 and len(idx) == len(t) + 1
    r = read_bundle(prefix, verify_crc=True)
    assert set(r) == set(t) and all(np.array_equal(r[k], t[k]) and r[k].shape == np.shape(t[k]) for k in t)
    sub = read_bundle(prefix, names=lambda n: n.startswith("decode/"))
    assert list(sub) == ["decode/final_layer_weights/bias"]
    with open(prefix + ".index", "r+b") as f:          # corrupt one byte of the first block -> CRC error
        f.seek(10)
        b = f.read(1)
        f.seek(10)
        f.write(bytes([b[0] ^ 0xFF]))
    with pytest.raises(ValueError):
        read_index(prefix + ".index")


def test_synthetic_tasks_schema_and_determinism():
    from mliis_b200.synthetic import SyntheticSegmentationTask, make_task_arrays
    iu8, mu8 = make_task_arrays(7, 4, 64)
    assert iu8.dtype == np.uint8 and iu8.shape == (4, 64, 64, 3) and set(np.unique(mu8)) <= {0, 255}
    frac = (mu8 == 255).mean(axis=(1, 2))
    assert np.all(frac >= 0.05) and np.all(frac <= 0.60)
    iu8b, mu8b = make_task_arrays(7, 4, 64)
    assert np.array_equal(iu8, iu8b) and np.array_equal(mu8, mu8b)
    t = SyntheticSegmentationTask(7, 4, 64)
    s = t.sample(None, 3)
    assert len(s) == 3 and s[0][0].shape == (64, 64, 3) and s[0][1].shape == (64, 64, 2)
    assert np.array_equal(s[0][1][..., 0] + s[0][1][..., 1], np.ones((64, 64), np.float32))     # one-hot
    assert s[0][0].max() <= 255 and s[0][0].min() >= 0 and t.name == "synthetic_0007"
    with pytest.raises(ValueError):
        t.sample(None, 5)


def test_model_surface_without_gpu():
    from mliis_b200.efficientlab import EfficientLab
    from mliis_b200.args import argument_parser, model_kwargs
    a = argument_parser().parse_args("--rsd 2 4 --l2 --loss_name bce_dice --image_size 224".split())
    m = EfficientLab(**model_kwargs(a))
    for attr in ["input_ph", "label_ph", "minimize_op", "predictions", "is_training_ph", "lr_ph",
                 "final_layer_dropout_rate_ph", "loss", "variables_initialized", "feature_extractor_name",
                 "final_layer_scope", "restore_model"]:
        assert hasattr(m, attr), attr
    assert m.final_layer_scope == "decode/final_layer_weights" and m.n_params == 2071714
    gv = m.global_variables()
    assert len(m.trainable_variables()) == 169 and len(gv) == 169 + 78 + 2 + 169
    assert gv[3].name.endswith("stem/tpu_batch_normalization/moving_mean")
    with pytest.raises(NotImplementedError):
        EfficientLab(rsd=[2, 4], feature_extractor_name="efficientnet-b3", learning_rate=1e-3, label_smoothing=0.0)
    with pytest.raises(NotImplementedError):
        EfficientLab(rsd=[2, 4], spatial_pyramid_pooling=True, learning_rate=1e-3, label_smoothing=0.0)


def test_native_crc32c_and_bundle_data_checksums(tmp_path):
    """mliis_crc32c (host code of the C library) against the byte-at-a-time Python routine and the CRC-32C check value;
    bundles carry per-tensor CRCs by default and a corrupted data shard is detected; Adam first-moment slots are written
    as zeros so that a strict TF restore finds every name."""
    from mliis_b200 import native
    from mliis_b200.checkpoint import _crc32c_py, crc32c, read_bundle, read_index, with_adam_m_slots, write_bundle
    rng = np.random.default_rng(0)
    lib = native.lib()
    assert lib.mliis_crc32c(b"123456789", 9, 0) == 0xE3069283
    for n in (0, 1, 7, 8, 9, 63, 255, 256, 257, 1000, 4099, 65539):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert crc32c(b) == _crc32c_py(b) == lib.mliis_crc32c(b, n, 0), n
        k = n // 3                                             # continuation from a running value
        assert lib.mliis_crc32c(b[k:], n - k, lib.mliis_crc32c(b[:k], k, 0)) == _crc32c_py(b), n
    t = with_adam_m_slots({"a/kernel": rng.standard_normal((3, 5)).astype("f4"),
                           "a/kernel/Adam_1": np.abs(rng.standard_normal((3, 5))).astype("f4"),
                           "beta2_power": np.float32(0.5)})
    assert set(t) == {"a/kernel", "a/kernel/Adam", "a/kernel/Adam_1", "beta2_power"} and not t["a/kernel/Adam"].any()
    prefix = str(tmp_path / "model.ckpt-3")
    write_bundle(prefix, t)
    assert all(e["crc32c"] != 0 for k, e in read_index(prefix + ".index").items() if k)
    r = read_bundle(prefix)
    assert all(np.array_equal(r[k], t[k]) for k in t)
    with open(prefix + ".data-00000-of-00001", "r+b") as f:     # flip one bit of the first tensor
        f.seek(5)
        b = f.read(1)
        f.seek(5)
        f.write(bytes([b[0] ^ 0x10]))
    with pytest.raises(ValueError):
        read_bundle(prefix)
