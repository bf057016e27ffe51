"""TensorFlow V2 checkpoint bundles without TensorFlow (restore: run_metasegnet.py:125-133; save: train.py:129-131).

Format [TF-ext, restated from the published TensorFlow sources tensor_bundle.{h,cc}, table format of LevelDB]:
  <dir>/checkpoint                      text: model_checkpoint_path: "model.ckpt-N"
  model.ckpt-N.index                    SSTable.  Blocks of prefix-compressed entries
                                          [shared varint][non_shared varint][value_len varint][key suffix][value]
                                        + restart array (uint32[] + count), each block followed by a 5-byte trailer
                                        (1 byte compression: 0 none / 1 snappy, masked CRC32C of block+type);
                                        48-byte footer = metaindex handle, index handle, padding,
                                        magic 0xdb4775248b80fb57.  Key "" -> BundleHeaderProto; key <variable
                                        name> -> BundleEntryProto{dtype, shape, shard_id, offset, size, crc32c}.
  model.ckpt-N.data-00000-of-00001      raw little-endian tensors.
The shipped EfficientLab-6-3_FOMAML-star tarball is absent from the reference mount (.MISSING_LARGE_BLOBS), so the
reader is verified against bundles produced by the writer below (tests/test_checkpoint.py) and the expected
variable names must be re-checked against the real index the day it is available.
"""
from __future__ import annotations

import os
import struct
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
DT_FLOAT = 1

# ---- CRC32C (Castagnoli), table driven -------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def _crc32c_py(data: bytes) -> int:
    """Byte-at-a-time CRC32C in Python (small buffers, and the cross-check of the native routine in the tests)."""
    t = _crc_table()
    c = 0xFFFFFFFF
    tl = t.tolist()
    for b in data:
        c = tl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data) -> int:
    """CRC-32C of a bytes-like object.  Buffers of 256 bytes and more go through the library's slicing-by-8 routine
    (mliis_crc32c, host code: ~1 GB/s, which is what makes per-tensor data CRCs affordable for 8 MB of weights)."""
    if len(data) < 256:
        return _crc32c_py(bytes(data))
    from . import native
    buf = data if isinstance(data, bytes) else bytes(data)
    return int(native.lib().mliis_crc32c(buf, len(buf), 0))


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def _unmask(m: int) -> int:
    rot = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format ----------------------------------------------------------------
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """field number -> list of raw values (ints for varint/fixed, bytes for length-delimited)."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(field: int, wt: int, payload: bytes) -> bytes:
    return _put_varint((field << 3) | wt) + payload


def _encode_shape(shape) -> bytes:
    out = b""
    for d in shape:
        dim = _field(1, 0, _put_varint(int(d)))
        out += _field(2, 2, _put_varint(len(dim)) + dim)
    return out


def _decode_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for d in _parse_proto(buf).get(2, []):
        dims.append(int(_parse_proto(d).get(1, [0])[0]))
    return tuple(dims)


# ---- snappy (decode only; TF bundles are normally written uncompressed) -----------------------------
def _snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


# ---- SSTable ------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if _unmask(stored) != crc32c(data[offset:offset + size + 1]):
            raise ValueError("index block CRC mismatch")
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy_decompress(raw)
    raise ValueError("unknown block compression %d" % ctype)


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_index(path: str, verify_crc: bool = True) -> Dict[str, dict]:
    """Parse model.ckpt-N.index -> {variable name: {dtype, shape, shard_id, offset, size, crc32c}} (+ '' header)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TF checkpoint index (bad SSTable magic)" % path)
    footer = data[-48:]
    pos = 0
    _, pos = _get_varint(footer, pos)      # metaindex handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    entries: Dict[str, dict] = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify_crc)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        for key, val in _block_entries(_read_block(data, off, size, verify_crc)):
            pr = _parse_proto(val)
            if key == b"":
                entries[""] = {"num_shards": pr.get(1, [1])[0], "endianness": pr.get(2, [0])[0]}
                continue
            entries[key.decode()] = {
                "dtype": pr.get(1, [0])[0], "shape": _decode_shape(pr[2][0]) if 2 in pr else (),
                "shard_id": pr.get(3, [0])[0], "offset": pr.get(4, [0])[0], "size": pr.get(5, [0])[0],
                "crc32c": pr.get(6, [0])[0], "sliced": 7 in pr}
    return entries


def read_bundle(prefix: str, names: Optional[Callable[[str], bool]] = None, verify_crc: bool = True
                ) -> Dict[str, np.ndarray]:
    """Load the float32 tensors of a V2 bundle `prefix` (= .../model.ckpt-N).  Per-tensor CRC-32Cs are verified as
    TensorFlow's BundleReader does; an entry whose stored CRC is 0 (bundles of older versions of this writer) is not."""
    index = read_index(prefix + ".index")
    header = index.pop("", {"num_shards": 1, "endianness": 0})
    if header.get("endianness", 0) != 0:
        raise ValueError("big-endian bundles are not supported")
    n_shards = header.get("num_shards", 1)
    shards = {}
    out = {}
    for name, e in index.items():
        if names is not None and not names(name):
            continue
        if e["dtype"] != DT_FLOAT:
            continue                       # e.g. int64 global_step
        if e["sliced"]:
            raise ValueError("partitioned variable %s is not supported" % name)
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, n_shards), dtype=np.uint8, mode="r")
        raw = bytes(shards[sid][e["offset"]:e["offset"] + e["size"]])
        if verify_crc and e["crc32c"] != 0 and _unmask(e["crc32c"]) != crc32c(raw):
            raise ValueError("tensor %s: CRC mismatch" % name)
        out[name] = np.frombuffer(raw, dtype="<f4").reshape(e["shape"]).copy()
    return out


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], with_data_crc: bool = True) -> None:
    """Write float32 tensors as a single-shard V2 bundle.  Index blocks and, by default, every tensor entry carry valid
    masked CRC-32Cs (TensorFlow's BundleReader::GetValue verifies the per-tensor one)."""
    names = sorted(tensors.keys())
    data = bytearray()
    entries = [(b"", _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) +
                _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]
    for n in names:
        a = np.asarray(tensors[n], dtype="<f4")
        raw = a.tobytes(order="C")
        shape = _encode_shape(a.shape)
        val = _field(1, 0, _put_varint(DT_FLOAT)) + _field(2, 2, _put_varint(len(shape)) + shape)
        val += _field(4, 0, _put_varint(len(data))) + _field(5, 0, _put_varint(len(raw)))
        val += _field(6, 5, struct.pack("<I", _mask(crc32c(raw)) if with_data_crc else 0))
        entries.append((n.encode(), val))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))

    out = bytearray()

    def emit_block(kvs, restart_interval=16) -> Tuple[int, int]:
        blk = bytearray()
        restarts, last = [], b""
        for i, (k, v) in enumerate(kvs):
            shared = 0
            if i % restart_interval == 0:
                restarts.append(len(blk))
            else:
                while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                    shared += 1
            blk += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
            last = k
        if not restarts:
            restarts = [0]
        for r in restarts:
            blk += struct.pack("<I", r)
        blk += struct.pack("<I", len(restarts))
        off = len(out)
        out.extend(blk)
        out.append(0)                                             # no compression
        out.extend(struct.pack("<I", _mask(crc32c(bytes(blk) + b"\x00"))))
        return off, len(blk)

    index_kvs = []
    chunk = 64
    for i in range(0, len(entries), chunk):
        part = entries[i:i + chunk]
        off, size = emit_block(part)
        index_kvs.append((part[-1][0] + b"\x00" if part[-1][0] == b"" else part[-1][0],
                          _put_varint(off) + _put_varint(size)))
    meta_off, meta_size = emit_block([])
    idx_off, idx_size = emit_block(index_kvs, restart_interval=1)
    footer = _put_varint(meta_off) + _put_varint(meta_size) + _put_varint(idx_off) + _put_varint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


# ---- glue to the model ----------------------------------------------------------------------------------
def model_tensors(model) -> Dict[str, np.ndarray]:
    """All global variables of the model as {TF name: array} (trainables, BN moving statistics, Adam slots)."""
    from .variables import VariableState
    sess = _SessionOf(model)
    variables = model.global_variables()
    vals = VariableState(sess, variables).export_variables()
    return {v.name: np.asarray(a, np.float32) for v, a in zip(variables, vals)}


def with_adam_m_slots(tensors: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """TF's AdamOptimizer also owns a first-moment slot `<var>/Adam` per variable.  With beta1 = 0 it is dead (m == g)
    and the engine does not keep it, but tf.train.Saver().restore is strict on names: write it as zeros next to every
    second-moment slot `<var>/Adam_1`."""
    out = dict(tensors)
    for name in tensors:
        if name.endswith("/Adam_1"):
            out.setdefault(name[:-2], np.zeros_like(tensors[name]))
    return out


class _SessionOf:
    def __init__(self, model):
        self.model = model


def restore_into_engine(model, ckpt_dir_or_prefix: str, keep: Optional[Callable[[str], bool]] = None,
                        strict: bool = True, warn_missing: bool = False) -> int:
    """Restore variables by name.  strict=True mirrors tf.train.Saver().restore: every global variable of the model
    must exist in the checkpoint (run_metasegnet.py:131-133) - Adam slots included when the model uses Adam.
    strict=False keeps the initial value of a variable the checkpoint lacks; warn_missing then says so loudly."""
    from .util import latest_checkpoint
    from .variables import VariableState
    prefix = ckpt_dir_or_prefix
    if os.path.isdir(prefix):
        prefix = latest_checkpoint(prefix)
    tensors = read_bundle(prefix)
    variables = [v for v in model.global_variables() if keep is None or keep(v.name)]
    missing = [v.name for v in variables if v.name not in tensors]
    if missing and strict:
        raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (prefix, len(missing), missing[:3]))
    if missing and warn_missing:
        import warnings
        warnings.warn("checkpoint %s lacks %d of the %d selected variables (e.g. %s): they keep their initial values; the "
                      "reference's Saver(var_dict).restore would have failed here (efficientlab.py:398-443)"
                      % (prefix, len(missing), len(variables), missing[:3]))
    present = [v for v in variables if v.name in tensors]
    for v in present:
        if tuple(tensors[v.name].shape) != tuple(v.shape):
            raise ValueError("shape mismatch for %s: checkpoint %s, model %s" % (v.name, tensors[v.name].shape, v.shape))
    VariableState(_SessionOf(model), present).import_variables([tensors[v.name] for v in present])
    return len(present)


class Saver:
    """tf.train.Saver stand-in: save(sess, path, global_step) / restore(sess, prefix), max_to_keep rotation and the
    `checkpoint` state file that utils.latest_checkpoint parses."""

    def __init__(self, model, max_to_keep: int = 2):
        self.model = model
        self.max_to_keep = max_to_keep
        self._kept: List[str] = []

    def save(self, sess, save_path: str, global_step: Optional[int] = None) -> str:
        prefix = save_path if global_step is None else "%s-%d" % (save_path, global_step)
        try:                               # one writer per job: the trainables are replicated, the files are shared
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
                return prefix
        except ImportError:
            pass
        write_bundle(prefix, with_adam_m_slots(model_tensors(self.model)))
        self._kept.append(prefix)
        while self.max_to_keep and len(self._kept) > self.max_to_keep:
            old = self._kept.pop(0)
            for suffix in (".index", ".data-00000-of-00001"):
                try:
                    os.remove(old + suffix)
                except OSError:
                    pass
        d = os.path.dirname(prefix)
        with open(os.path.join(d, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
            for p in self._kept:
                f.write('all_model_checkpoint_paths: "%s"\n' % os.path.basename(p))
        return prefix

    def restore(self, sess, prefix: str) -> None:
        restore_into_engine(self.model, prefix, strict=True)
        self.model.variables_initialized = True
