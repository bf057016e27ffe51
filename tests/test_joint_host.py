"""Host side of the joint-training path on CPU: shard discovery, sparse example loading from both
shard layouts, batching / rank slicing, schedules, and the data-parallel gradient average over gloo (world size 2)."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mliis_b200 import fss1000, joint_train as jt, tfrecord
from mliis_b200.synthetic import make_task_arrays


def _write_class_shards(root, names, n=4, size=16):
    out = {}
    for t, name in enumerate(names):
        iu8, mu8 = make_task_arrays(t, n, size)
        fss1000.write_task_shard(os.path.join(root, name + ".tfrecord.gzip"), iu8, mu8)
        out[name] = (iu8, mu8)
    return out


def test_shard_discovery_by_name(tmp_path):
    for n in ("train-000", "train-001", "test-000"):
        open(str(tmp_path / (n + ".tfrecord.gzip")), "wb").close()
    open(str(tmp_path / "notes.txt"), "w").close()
    tr, te = jt.get_train_test_shards_from_dir(str(tmp_path))
    assert sorted(map(os.path.basename, tr)) == ["train-000.tfrecord.gzip", "train-001.tfrecord.gzip"]
    assert sorted(map(os.path.basename, te)) == ["test-000.tfrecord.gzip"]
    # a 'val' shard is neither 'train' nor 'test' by name: the reference's own asserts (joint_train.py:131-134) only
    # accept it together with --test_on_val_set
    open(str(tmp_path / "val-000.tfrecord.gzip"), "wb").close()
    with pytest.raises(AssertionError):
        jt.get_train_test_shards_from_dir(str(tmp_path))
    tr, te = jt.get_train_test_shards_from_dir(str(tmp_path), test_on_val_set=True)
    assert sorted(map(os.path.basename, tr)) == ["train-000.tfrecord.gzip", "train-001.tfrecord.gzip"]
    assert sorted(map(os.path.basename, te)) == ["val-000.tfrecord.gzip"]


def test_sparse_loading_from_class_shards_and_dense_joint_shards(tmp_path):
    names = ["bus", "ant", "cat"]                       # sorted: ant=1, bus=2, cat=3
    arrays = _write_class_shards(str(tmp_path), names)
    classes = sorted(names)
    paths = [os.path.join(str(tmp_path), n + ".tfrecord.gzip") for n in names]
    data = jt.load_sparse_shards(paths, 16, classes)
    assert len(data) == 12 and data.n_classes == 3
    assert data.class_ids.tolist() == [2] * 4 + [1] * 4 + [3] * 4
    np.testing.assert_array_equal(data.images[:4], arrays["bus"][0])
    np.testing.assert_array_equal(data.masks[4:8], arrays["ant"][1])
    # the reference's dense joint shard: mask bytes [S,S,C] with channel 0 = 255 - mask (one_hot_encode, writer :140-158)
    dense_dir = tmp_path / "dense"
    dense_dir.mkdir()
    payloads = []
    for r in range(len(data)):
        m = data.masks[r]
        d = np.zeros((16, 16, 4), np.uint8)
        d[:, :, 0] = 255 - m
        d[:, :, data.class_ids[r]] = m
        payloads.append(tfrecord.encode_example({"image": data.images[r].tobytes(), "mask": d.tobytes()}))
    tfrecord.write_tfrecords(str(dense_dir / "train-0.tfrecord.gzip"), payloads)
    again = jt.load_sparse_shards([str(dense_dir / "train-0.tfrecord.gzip")], 16, classes)
    np.testing.assert_array_equal(again.images, data.images)
    np.testing.assert_array_equal(again.masks, data.masks)
    np.testing.assert_array_equal(again.class_ids, data.class_ids)
    lab = data.dense_labels([0, 5])
    assert lab.shape == (2, 16, 16, 4) and np.all(lab.sum(-1) == 1.0) and lab[1, :, :, 2].sum() == 0
    with pytest.raises(tfrecord.TFRecordError):
        jt.load_sparse_shards(paths, 16, ["ant", "cat"])            # 'bus' has no class id


def test_batcher_is_seeded_and_ranks_partition_the_global_batch():
    iu8, mu8 = make_task_arrays(0, 10, 8)
    data = jt.SparseSegmentationData(iu8, mu8, np.arange(10, dtype=np.int32) % 3 + 1, 3)
    whole = jt.SparseBatcher(data, 4, seed=5)
    r0, r1 = jt.SparseBatcher(data, 4, seed=5, rank=0, world=2), jt.SparseBatcher(data, 4, seed=5, rank=1, world=2)
    seen = []
    for _ in range(6):
        rows = whole.next_rows()
        a, b = r0.next_rows(), r1.next_rows()
        assert np.array_equal(np.concatenate([a, b]), rows)
        seen.extend(rows.tolist())
    assert sorted(seen[:10]) == list(range(10))                    # an epoch visits every example once
    x, m, c = jt.SparseBatcher(data, 4, seed=5).next_batch()
    assert x.dtype == np.float32 and m.dtype == np.float32 and c.dtype == np.int32
    assert x.shape == (4, 8, 8, 3) and m.shape == (4, 8, 8) and set(np.unique(m)) <= {0.0, 1.0}
    with pytest.raises(ValueError):
        jt.SparseBatcher(data, 5, world=2)
    # augmentation keeps the mask binary-valued for the label-preserving transforms and the shapes intact
    random.seed(0)
    np.random.seed(0)
    xa, ma, _ = jt.SparseBatcher(data, 4, seed=5, augmenter=jt.make_augmenter()).next_batch()
    assert xa.shape == x.shape and ma.shape == m.shape and set(np.unique(ma)) <= {0.0, 1.0}
    assert not np.array_equal(xa, x)


def test_schedule_and_model_kwargs():
    assert jt.linear_lr(0, 10, 0.005, 5e-7) == 0.005
    assert abs(jt.linear_lr(5, 10, 0.005, 5e-7) - 0.5 * (0.005 + 5e-7)) < 1e-12
    a = jt.parse_args(["--data_dir", "/d", "--rsd", "2", "4", "--sgd", "--l2", "--loss_name", "cross_entropy",
                       "--seperate_background_channel", "--batch_size", "32"])
    kw = jt.get_model_kwargs(a)
    assert kw["rsd"] == [2, 4] and kw["optimizer"] == "sgd" and kw["dice"] is False and kw["l2"] is True
    assert kw["n_rows"] == 224 and a.batch_size == 32
    # the reference's own command lines (joint_train.py:4-6) parse
    jt.parse_args("--seperate_background_channel --data_dir d --augment --epochs 10 --steps_per_epoch 2 --batch_size 3 "
                  "--val_batches 2 --sgd --l2 --final_layer_dropout_rate 0.2 --rsd 2 "
                  "--restore_efficient_net_weights_from models/efficientnet/efficientnet-b0".split())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(8, dtype=torch.float32) * (rank + 1)
    d, r, w = jt._dist()
    jt.average_gradients(g, d, w)
    out[rank] = g.tolist()
    dist.destroy_process_group()


def test_data_parallel_gradient_average_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    expect = (torch.arange(8, dtype=torch.float32) * 1.5).tolist()
    assert out[0] == expect and out[1] == expect


def test_shard_layout_detection(tmp_path):
    """per-class shards (`<class>.tfrecord.gzip`) are not split by 'train'/'test' substrings."""
    from mliis_b200.joint_train import get_train_test_shards_from_dir
    d = tmp_path / "per_class"
    d.mkdir()
    for n in ("contest", "train_station", "bus", "eagle"):
        (d / (n + ".tfrecord.gzip")).write_bytes(b"")
    tr, te = get_train_test_shards_from_dir(str(d), test_ids=["bus"])
    assert sorted(os.path.basename(p) for p in te) == ["bus.tfrecord.gzip"]
    assert sorted(os.path.basename(p) for p in tr) == ["contest.tfrecord.gzip", "eagle.tfrecord.gzip",
                                                       "train_station.tfrecord.gzip"]
    tr, te = get_train_test_shards_from_dir(str(d))
    assert len(tr) == 4 and te == []
    d2 = tmp_path / "dense"
    d2.mkdir()
    for n in ("train_0", "train_1", "test_0"):
        (d2 / (n + ".tfrecord.gzip")).write_bytes(b"")
    tr, te = get_train_test_shards_from_dir(str(d2))
    assert sorted(os.path.basename(p) for p in tr) == ["train_0.tfrecord.gzip", "train_1.tfrecord.gzip"]
    assert sorted(os.path.basename(p) for p in te) == ["test_0.tfrecord.gzip"]
