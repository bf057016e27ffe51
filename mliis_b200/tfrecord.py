"""Gzip-TFRecord reader / writer for the FSS-1000 shards, without TensorFlow.

Replaces, on the host side of the hot path:
  * ``tf.data.TFRecordDataset(compression_type="GZIP")`` + ``parse_example``       data/input_fn.py:28-65, :103-115
  * ``tf.python_io.tf_record_iterator`` in ``count_examples_in_tfrecords``          utils/util.py:24-33
  * the writer ``write_tfrecord`` / ``make_example``               data/fss_1000_image_to_tfrecord.py:99-134, :137-160

Formats [TF-ext]:
  TFRecord framing   u64 length (LE) | u32 masked_crc32c(length bytes) | payload | u32 masked_crc32c(payload)
                     mask(c) = ((c >> 15) | (c << 17)) + 0xa282ead8  (mod 2^32), crc = CRC-32C (Castagnoli)
  tf.train.Example   Example{1: Features{1: map<string, Feature>}},  map entry {1: key, 2: Feature},
                     Feature{1: BytesList{1: repeated bytes} | 2: FloatList | 3: Int64List}
  FSS-1000 record    bytes features ``image`` (S*S*3 uint8, HWC) and ``mask`` (S*S uint8, positive class = 255)
"""
from __future__ import annotations

import glob
import gzip
import struct
from typing import Dict, Iterable, Iterator, List, Sequence, Tuple, Union

import numpy as np

from .checkpoint import _field, _get_varint, _mask, _parse_proto, _put_varint, crc32c


class TFRecordError(ValueError):
    pass


def _ld(field: int, payload: bytes) -> bytes:
    """length-delimited protobuf field"""
    return _field(field, 2, _put_varint(len(payload)) + payload)


def _open(path: str, mode: str, compression: str):
    c = (compression or "").upper()
    if c == "GZIP":
        return gzip.open(path, mode)
    if c in ("", "NONE"):
        return open(path, mode)
    raise TFRecordError("unsupported TFRecord compression type %r" % compression)


def read_tfrecords(path: str, compression: str = "GZIP", verify_crc: bool = True) -> Iterator[bytes]:
    """Yields the payload of every record of one shard, in file order."""
    with _open(path, "rb", compression) as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) != 12:
                raise TFRecordError("%s: truncated record header" % path)
            (length,), (len_crc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify_crc and _mask(crc32c(head[:8])) != len_crc:
                raise TFRecordError("%s: corrupted record length" % path)
            body = f.read(length + 4)
            if len(body) != length + 4:
                raise TFRecordError("%s: truncated record" % path)
            payload = body[:length]
            if verify_crc and _mask(crc32c(payload)) != struct.unpack("<I", body[length:])[0]:
                raise TFRecordError("%s: corrupted record payload" % path)
            yield payload


def write_tfrecords(path: str, payloads: Iterable[bytes], compression: str = "GZIP") -> int:
    n = 0
    with _open(path, "wb", compression) as f:
        for p in payloads:
            head = struct.pack("<Q", len(p))
            f.write(head)
            f.write(struct.pack("<I", _mask(crc32c(head))))
            f.write(p)
            f.write(struct.pack("<I", _mask(crc32c(p))))
            n += 1
    return n


# ---------------------------------------------------------------------------------------------------
# tf.train.Example (bytes / int64 / float features)
# ---------------------------------------------------------------------------------------------------
def decode_example(payload: bytes) -> Dict[str, list]:
    """{feature name: list of bytes | list of int | list of float}."""
    out: Dict[str, list] = {}
    ex = _parse_proto(payload)
    for features in ex.get(1, []):
        for entry in _parse_proto(features).get(1, []):
            e = _parse_proto(entry)
            key = e[1][0].decode("utf-8")
            feat = _parse_proto(e[2][0]) if 2 in e else {}
            if 1 in feat:        # BytesList
                out[key] = list(_parse_proto(feat[1][0]).get(1, []))
            elif 3 in feat:      # Int64List (packed or not)
                vals: List[int] = []
                for v in _parse_proto(feat[3][0]).get(1, []):
                    if isinstance(v, (bytes, bytearray)):
                        pos = 0
                        while pos < len(v):
                            x, pos = _get_varint(v, pos)
                            vals.append(x - (1 << 64) if x >= (1 << 63) else x)
                    else:
                        vals.append(v - (1 << 64) if v >= (1 << 63) else v)
                out[key] = vals
            elif 2 in feat:      # FloatList (packed or not)
                fl: List[float] = []
                for v in _parse_proto(feat[2][0]).get(1, []):
                    if isinstance(v, (bytes, bytearray)) and len(v) % 4 == 0:
                        fl.extend(struct.unpack("<%df" % (len(v) // 4), v))
                out[key] = fl
            else:
                out[key] = []
    return out


def encode_example(features: Dict[str, Union[bytes, Sequence[bytes]]]) -> bytes:
    """Bytes features only (what the FSS-1000 writer emits).  Keys are written in sorted order, like protobuf's
    deterministic map serialisation."""
    body = b""
    for key in sorted(features):
        v = features[key]
        values = [v] if isinstance(v, (bytes, bytearray)) else list(v)
        blist = b"".join(_ld(1, bytes(x)) for x in values)
        feature = _ld(1, blist)
        entry = _ld(1, key.encode("utf-8")) + _ld(2, feature)
        body += _ld(1, entry)
    return _ld(1, body)


# ---------------------------------------------------------------------------------------------------
# FSS-1000 records
# ---------------------------------------------------------------------------------------------------
def parse_example(payload: bytes, image_width: int, scale_to_0_1: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """data/input_fn.py:28-65: (image f32 [S,S,3] in 0..255, mask f32 [S,S,2] = stack([255-m, m]) / 255)."""
    f = decode_example(payload)
    if "image" not in f or "mask" not in f or not f["image"] or not f["mask"]:
        raise TFRecordError("record lacks the 'image' / 'mask' bytes features")
    img = np.frombuffer(f["image"][0], np.uint8)
    msk = np.frombuffer(f["mask"][0], np.uint8)
    if img.size != image_width * image_width * 3 or msk.size != image_width * image_width:
        raise TFRecordError("record is not %dx%d (image %d bytes, mask %d bytes)"
                            % (image_width, image_width, img.size, msk.size))
    image = img.reshape(image_width, image_width, 3).astype(np.float32)
    if scale_to_0_1:
        image /= np.float32(255.0)
    m = msk.reshape(image_width, image_width)
    mask = np.stack([255 - m, m], axis=2).astype(np.float32) / np.float32(255.0)
    return image, mask


def make_example(image_u8: np.ndarray, mask_u8: np.ndarray) -> bytes:
    """data/fss_1000_image_to_tfrecord.py:117-134 (mask: first channel only, positive class 255)."""
    image_u8 = np.ascontiguousarray(image_u8, np.uint8)
    mask_u8 = np.asarray(mask_u8, np.uint8)
    if mask_u8.ndim > 2:
        mask_u8 = mask_u8[:, :, 0]
    return encode_example({"image": image_u8.tobytes(), "mask": np.ascontiguousarray(mask_u8).tobytes()})


def expand_paths(paths: Union[str, Sequence[str]]) -> List[str]:
    """A path, a glob or a list of either -> sorted list of shard files (``Dataset.list_files`` shuffles; with one
    shard per task, the reference's layout, the order is the same)."""
    if isinstance(paths, str):
        paths = [paths]
    out: List[str] = []
    for p in paths:
        hits = sorted(glob.glob(p))
        out.extend(hits if hits else [])
    seen, uniq = set(), []
    for p in out:
        if p not in seen:
            seen.add(p)
            uniq.append(p)
    return uniq


def count_examples_in_tfrecords(paths: Union[str, Sequence[str]], compression: str = "GZIP") -> int:
    """utils/util.py:24-33."""
    if isinstance(paths, str):
        paths = [paths]
    n = 0
    for fn in paths:
        for _ in read_tfrecords(fn, compression, verify_crc=False):
            n += 1
    return n


def load_examples(paths: Union[str, Sequence[str]], image_width: int, limit: int = None,
                  compression: str = "GZIP") -> Tuple[np.ndarray, np.ndarray]:
    """First ``limit`` records (file order) of the shards as (images f32 [n,S,S,3], masks f32 [n,S,S,2])."""
    images, masks = [], []
    for fn in expand_paths(paths):
        for payload in read_tfrecords(fn, compression):
            if limit is not None and len(images) >= limit:
                break
            im, mk = parse_example(payload, image_width)
            images.append(im)
            masks.append(mk)
    if not images:
        return (np.zeros((0, image_width, image_width, 3), np.float32),
                np.zeros((0, image_width, image_width, 2), np.float32))
    return np.stack(images), np.stack(masks)
