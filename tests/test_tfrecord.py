"""Gzip-TFRecord / tf.train.Example reader and the FSS-1000 task readers (CPU only).

The wire format is pinned independently of our own encoder: records are cross-checked against the protobuf
runtime with the tf.train.Example schema (feature.proto / example.proto [TF-ext]) declared programmatically,
and the framing CRC against the CRC-32C known answer.
"""
import gzip
import os
import random
import struct

import numpy as np
import pytest

from mliis_b200 import fss1000, tfrecord
from mliis_b200.checkpoint import _mask, crc32c
from mliis_b200.synthetic import make_task_arrays, parse_records


def _example_classes():
    """tf.train.Example message classes built from the published schema (no TensorFlow)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="mliis_test_example.proto", package="tfx", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, typ, label=T.LABEL_OPTIONAL, type_name=None, oneof=None, packed=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, typ, label
        if type_name:
            f.type_name = type_name
        if oneof is not None:
            f.oneof_index = oneof
        if packed is not None:
            f.options.packed = packed
        return f

    field(msg("BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(msg("FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    field(msg("Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    feat = msg("Feature")
    feat.oneof_decl.add().name = "kind"
    field(feat, "bytes_list", 1, T.TYPE_MESSAGE, type_name=".tfx.BytesList", oneof=0)
    field(feat, "float_list", 2, T.TYPE_MESSAGE, type_name=".tfx.FloatList", oneof=0)
    field(feat, "int64_list", 3, T.TYPE_MESSAGE, type_name=".tfx.Int64List", oneof=0)
    feats = msg("Features")
    entry = feats.nested_type.add()
    entry.name = "FeatureEntry"
    entry.options.map_entry = True
    field(entry, "key", 1, T.TYPE_STRING)
    field(entry, "value", 2, T.TYPE_MESSAGE, type_name=".tfx.Feature")
    field(feats, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name=".tfx.Features.FeatureEntry")
    field(msg("Example"), "features", 1, T.TYPE_MESSAGE, type_name=".tfx.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tfx.Example"))


def test_crc32c_known_answer_and_mask():
    assert crc32c(b"123456789") == 0xE3069283            # CRC-32C check value (RFC 3720 B.4)
    c = crc32c(b"123456789")
    assert _mask(c) == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_format_against_protobuf_runtime():
    Example = _example_classes()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (6, 6, 3), dtype=np.uint8)
    msk = (rng.random((6, 6)) > 0.5).astype(np.uint8) * 255
    ours = tfrecord.make_example(img, msk)
    # protobuf parses our bytes
    ex = Example.FromString(ours)
    assert ex.features.feature["image"].bytes_list.value[0] == img.tobytes()
    assert ex.features.feature["mask"].bytes_list.value[0] == msk.tobytes()
    # we parse protobuf's bytes (incl. int64 / float features, packed)
    ref = Example()
    ref.features.feature["image"].bytes_list.value.append(img.tobytes())
    ref.features.feature["mask"].bytes_list.value.append(msk.tobytes())
    ref.features.feature["label"].int64_list.value.extend([3, -7, 1 << 40])
    ref.features.feature["score"].float_list.value.extend([0.5, -2.0])
    got = tfrecord.decode_example(ref.SerializeToString())
    assert got["image"] == [img.tobytes()] and got["mask"] == [msk.tobytes()]
    assert got["label"] == [3, -7, 1 << 40]
    assert got["score"] == [0.5, -2.0]
    image, mask = tfrecord.parse_example(ref.SerializeToString(), 6)
    assert image.dtype == np.float32 and mask.dtype == np.float32
    np.testing.assert_array_equal(image, img.astype(np.float32))
    np.testing.assert_array_equal(mask[..., 1], msk.astype(np.float32) / np.float32(255))
    np.testing.assert_array_equal(mask[..., 0], (255 - msk).astype(np.float32) / np.float32(255))


def test_record_framing_round_trip_and_corruption(tmp_path):
    payloads = [b"", b"a", os.urandom(1000), b"x" * 70000]
    for comp in ("GZIP", ""):
        p = str(tmp_path / ("r." + (comp or "raw")))
        assert tfrecord.write_tfrecords(p, payloads, compression=comp) == 4
        assert list(tfrecord.read_tfrecords(p, compression=comp)) == payloads
    # framing layout of the first non-empty record (uncompressed): len | crc(len) | payload | crc(payload)
    raw = open(str(tmp_path / "r.raw"), "rb").read()
    assert raw[:8] == struct.pack("<Q", 0)
    off = 16
    assert raw[off:off + 8] == struct.pack("<Q", 1)
    assert struct.unpack("<I", raw[off + 8:off + 12])[0] == _mask(crc32c(struct.pack("<Q", 1)))
    assert raw[off + 12:off + 13] == b"a"
    assert struct.unpack("<I", raw[off + 13:off + 17])[0] == _mask(crc32c(b"a"))
    # flip one payload byte
    bad = bytearray(raw)
    bad[off + 12] ^= 0x40
    pb = str(tmp_path / "bad.raw")
    open(pb, "wb").write(bytes(bad))
    with pytest.raises(tfrecord.TFRecordError):
        list(tfrecord.read_tfrecords(pb, compression=""))
    # truncated file
    open(pb, "wb").write(raw[:-3])
    with pytest.raises(tfrecord.TFRecordError):
        list(tfrecord.read_tfrecords(pb, compression=""))
    with pytest.raises(tfrecord.TFRecordError):
        tfrecord.parse_example(tfrecord.encode_example({"image": b"abc"}), 4)


def _write_dataset(root, names, n_examples=7, size=16):
    arrays = {}
    for t, name in enumerate(names):
        iu8, mu8 = make_task_arrays(t, n_examples, size)
        fss1000.write_task_shard(os.path.join(root, name + ".tfrecord.gzip"), iu8, mu8)
        arrays[name] = (iu8, mu8)
    return arrays


def test_read_fss_1000_dataset_and_task_contract(tmp_path):
    names = ["ab_wheel", "bus", "crab", "dart", "eagle", "fox"]
    arrays = _write_dataset(str(tmp_path), names)
    open(str(tmp_path / "fss_test_set.txt"), "w").write("bus\neagle\nnot_present\n")
    tr, va, te, trn, van, ten = fss1000.read_fss_1000_dataset(str(tmp_path), num_val_tasks=1, image_size=16)
    assert sorted(t.name for t in te) == ["bus.tfrecord.gzip", "eagle.tfrecord.gzip"] and sorted(ten) == sorted(
        t.name for t in te)
    # reproducible val split: last of the sorted remaining shards
    assert [t.name for t in va] == ["fox.tfrecord.gzip"] and van == ["fox.tfrecord.gzip"]
    assert sorted(t.name for t in tr) == ["ab_wheel.tfrecord.gzip", "crab.tfrecord.gzip", "dart.tfrecord.gzip"]
    task = [t for t in te if t.name.startswith("bus")][0]
    assert task.batch_size == 7
    got = task.sample(None, 3)
    exp_i, exp_m = parse_records(*arrays["bus"])
    assert len(got) == 3
    for k in range(3):                                   # the FIRST n records, in file order
        np.testing.assert_array_equal(got[k][0], exp_i[k])
        np.testing.assert_array_equal(got[k][1], exp_m[k])
    ai, am = task.arrays()
    np.testing.assert_array_equal(ai, exp_i)
    np.testing.assert_array_equal(am, exp_m)
    with pytest.raises(ValueError):
        task.sample(None, 8)
    # explicit id list and explicit file path behave the same
    a = fss1000.read_fss_1000_dataset(str(tmp_path), test_task_ids=["bus", "eagle"], image_size=16)
    b = fss1000.read_fss_1000_dataset(str(tmp_path), test_task_ids=str(tmp_path / "fss_test_set.txt"), image_size=16)
    assert sorted(a[5]) == sorted(b[5]) == sorted(ten)


def test_random_split_consumes_the_random_stream_like_the_reference(tmp_path):
    names = ["t%02d" % i for i in range(9)]
    _write_dataset(str(tmp_path), names, n_examples=2, size=8)
    random.seed(3)
    tr, _, te, _, _, _ = fss1000.read_fss_1000_dataset(str(tmp_path), num_test_tasks=4, test_task_ids=None,
                                                       image_size=8)
    after = random.random()
    # data/fss_1000_utils.py:7-19: one random.shuffle of the globbed list, then pop() n_test times
    random.seed(3)
    shards = fss1000.get_fss_tasks(str(tmp_path))
    random.shuffle(shards)
    exp_test = [shards.pop() for _ in range(4)]
    assert [os.path.join(str(tmp_path), t.name) for t in te] == exp_test
    assert random.random() == after
    assert len(tr) == 5 and not set(t.name for t in tr) & set(t.name for t in te)


def test_fp_k_shot_tasks_merge_synonym_shards(tmp_path):
    names = ["aeroplane", "airliner", "bus", "motorbike", "potted_plant", "television", "zebra"]
    _write_dataset(str(tmp_path), names, n_examples=3, size=8)
    tasks, tnames = fss1000.read_fp_k_shot_dataset(
        str(tmp_path), all_task_names=[["airliner", "aeroplane"], ["bus"], ["potted plant"]], image_size=8)
    assert tnames == ["airliner", "bus", "pottedplant"]
    assert [t.batch_size for t in tasks] == [6, 3, 0]     # "potted plant" -> "pottedplant" matches no shard name
    assert tasks[0].arrays()[0].shape == (6, 8, 8, 3)
    assert len(tasks[1].sample(None, 2)) == 2


def test_count_examples_matches_gzip_stream(tmp_path):
    _write_dataset(str(tmp_path), ["a", "b"], n_examples=5, size=8)
    paths = sorted(fss1000.get_fss_tasks(str(tmp_path)))
    assert tfrecord.count_examples_in_tfrecords(paths) == 10
    raw = gzip.open(paths[0], "rb").read()
    assert len(raw) == 5 * (12 + 4 + len(tfrecord.make_example(np.zeros((8, 8, 3), np.uint8), np.zeros((8, 8), np.uint8))))


def test_image_folders_to_shards_tool(tmp_path):
    """tools/fss1000_to_tfrecords.py (the reference's data/fss_1000_image_to_tfrecord.py without TF / imageio)."""
    import importlib.util
    from PIL import Image
    spec = importlib.util.spec_from_file_location(
        "fss_tool", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                                 "fss1000_to_tfrecords.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    src, dst = tmp_path / "fewshot_data", tmp_path / "shards"
    rng = np.random.default_rng(0)
    truth = {}
    for cls in ("ab_wheel", "zebra"):
        (src / cls).mkdir(parents=True)
        for k in range(1, 4):
            size = 16 if not (cls == "zebra" and k == 3) else 12          # one wrongly sized pair is skipped
            img = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
            msk = (rng.random((size, size)) > 0.5).astype(np.uint8) * 255
            Image.fromarray(img).save(str(src / cls / ("%d.png" % k)).replace(".png", ".bmp"))   # lossless stand-in
            os.rename(str(src / cls / ("%d.bmp" % k)), str(src / cls / ("%d.jpg" % k)))
            Image.fromarray(np.stack([msk] * 3, axis=2)).save(str(src / cls / ("%d.png" % k)))
            if size == 16:
                truth[(cls, img.tobytes())] = msk
    random.seed(0)
    tool.main(["prog", "--input_dir", str(src), "--tfrecord_dir", str(dst), "--image_dims", "16"])
    tr, _, te, _, _, _ = fss1000.read_fss_1000_dataset(str(dst), test_task_ids=["zebra"], image_size=16)
    assert [t.name for t in tr] == ["ab_wheel.tfrecord.gzip"] and [t.batch_size for t in tr] == [3]
    assert [t.name for t in te] == ["zebra.tfrecord.gzip"] and [t.batch_size for t in te] == [2]
    for task, cls in ((tr[0], "ab_wheel"), (te[0], "zebra")):
        images, masks = task.arrays()
        for im, mk in zip(images, masks):
            want = truth[(cls, im.astype(np.uint8).tobytes())]               # pairs stay together
            np.testing.assert_array_equal(mk[..., 1], want.astype(np.float32) / np.float32(255))


def test_missing_official_test_split_is_an_error(tmp_path):
    """never fall back silently to a random split when the official FSS-1000 list is absent."""
    _write_dataset(str(tmp_path), ["a", "b", "c"], n_examples=2, size=8)
    with pytest.raises(FileNotFoundError):
        fss1000.read_fss_1000_dataset(str(tmp_path), image_size=8)
    tr, _, te, _, _, _ = fss1000.read_fss_1000_dataset(str(tmp_path), num_test_tasks=1, test_task_ids=None, image_size=8)
    assert len(tr) == 2 and len(te) == 1
