"""Reader (and a minimal writer) of TensorFlow "tensor bundle" checkpoints without TensorFlow.

The reference trains with ``tf.estimator`` and warm-starts from ``model.ckpt-N`` files
(``code/utils/framework_operations.py:126-129, 218-221``).  The engine keeps every variable under the
reference's name and in the Keras layout, so an existing IGNNITION checkpoint maps 1:1: this module
parses the two files of a bundle,

* ``<prefix>.index``: a leveldb-style sorted table (blocks of prefix-compressed key / value entries,
  an index block, a 48-byte footer with magic 0xdb4775248b80fb57) whose values are ``BundleEntryProto``
  messages (dtype, shape, shard, offset, size) -- the empty key holds the ``BundleHeaderProto``;
* ``<prefix>.data-SSSSS-of-NNNNN``: the raw little-endian tensor bytes,

restated from the published format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table).
No TensorFlow is installed here, so the parser is checked against files written by ``write`` below (the
same format, uncompressed blocks) and NOT against a TF-written file: parity of this reader is unpinned.
"""

from __future__ import annotations

import glob
import os
import re
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

MAGIC = 0xdb4775248b80fb57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ------------------------------------------------------------------ varints / protobuf
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _proto_fields(buf: bytes):
    """(field number, wire type, value) of one protobuf message; length-delimited values as bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise RuntimeError("IGNNITION: unsupported protobuf wire type %d in a checkpoint entry" % wt)
        yield field, wt, v


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard": 0, "offset": 0, "size": 0, "sliced": False}
    for field, _, v in _proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:                       # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
            for f2, _, d in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, s in _proto_fields(d):
                        if f3 == 1:
                            size = s
                    e["shape"].append(size)
        elif field == 3:
            e["shard"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 7:
            e["sliced"] = True
    return e


# ------------------------------------------------------------------ snappy (index blocks may be compressed)
def _snappy(buf: bytes) -> bytes:
    n, pos = _varint(buf, 0)
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
        raise RuntimeError("IGNNITION: corrupt snappy block in the checkpoint index")
    return bytes(out)


# ------------------------------------------------------------------ table
def _block(data: bytes, offset: int, size: int) -> bytes:
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy(raw)
    raise RuntimeError("IGNNITION: unknown block compression %d in the checkpoint index" % ctype)


def _block_entries(block: bytes):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(path: str) -> Dict[str, dict]:
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != MAGIC:
        raise RuntimeError("IGNNITION: %s is not a TensorFlow checkpoint index (bad magic)" % path)
    footer = data[-48:]
    _, pos = _varint(footer, 0)          # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    entries: Dict[str, dict] = {}
    for _, handle in _block_entries(_block(data, ioff, isize)):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        for key, value in _block_entries(_block(data, boff, bsize)):
            if key == b"":
                continue                      # BundleHeaderProto
            entries[key.decode()] = _parse_entry(value)
    return entries


def is_tf_checkpoint(path: str) -> bool:
    return not str(path).endswith(".npz") and os.path.exists(str(path) + ".index")


def latest(directory: str) -> Optional[str]:
    """Prefix of the newest ``model.ckpt-N`` bundle of a directory (by N)."""
    best, best_step = None, -1
    for f in glob.glob(os.path.join(directory, "*.index")):
        m = re.search(r"-(\d+)\.index$", f)
        step = int(m.group(1)) if m else 0
        if step > best_step:
            best, best_step = f[:-len(".index")], step
    return best


def read(prefix: str) -> Dict[str, np.ndarray]:
    """Every full (unsliced) tensor of the bundle, by variable name."""
    entries = read_index(prefix + ".index")
    shards = sorted(glob.glob(prefix + ".data-*"))
    n_shards = len(shards)
    files: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in entries.items():
        if e["sliced"] or e["dtype"] not in DTYPES:
            continue
        if e["shard"] not in files:
            path = "%s.data-%05d-of-%05d" % (prefix, e["shard"], n_shards)
            if not os.path.exists(path):
                raise RuntimeError("IGNNITION: checkpoint shard %s is missing" % path)
            files[e["shard"]] = open(path, "rb").read()
        raw = files[e["shard"]][e["offset"]:e["offset"] + e["size"]]
        out[name] = np.frombuffer(raw, dtype=np.dtype(DTYPES[e["dtype"]]).newbyteorder("<")).reshape(e["shape"]).copy()
    return out


def read_for_engine(prefix: str, engine) -> Tuple[Dict[str, np.ndarray], int]:
    """The engine's variables out of a reference checkpoint: a stored variable matches ``name`` when its key,
    with any ``:0`` and leading scope removed, ends with ``name`` and has the same shape; optimizer slots
    (``.../Adam``, ``.../Adam_1``) are skipped -- the reference warm-starts weights only (:129)."""
    stored = read(prefix)
    keys = [k for k in stored if not re.search(r"/(Adam|Adam_1|Momentum|RMSProp(_1)?)$", k)]
    weights, missing = {}, []
    for name, (_, shape) in engine.param_table.items():
        hit = [k for k in keys if (k == name or k.endswith("/" + name)) and tuple(stored[k].shape) == tuple(shape)]
        if len(hit) == 1:
            weights[name] = stored[hit[0]].astype(np.float32)
        else:
            missing.append(name)
    if missing:
        raise RuntimeError("IGNNITION: the checkpoint %s does not hold (or holds ambiguously) the variables %s; it has %s"
                           % (prefix, missing[:6], sorted(keys)[:12]))
    step = int(stored["global_step"]) if "global_step" in stored else 0
    return weights, step


# ------------------------------------------------------------------ writer (uncompressed, one shard)
def _entry_proto(dtype_id: int, shape, offset: int, size: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype_id)
    out += b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    return out


def _build_block(items: List[Tuple[bytes, bytes]]) -> bytes:
    body = bytearray()
    for key, value in items:                       # restart interval 1: no key sharing
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(value)) + key + value
    offs, pos = [], 0
    for key, value in items:
        offs.append(pos)
        pos += len(_put_varint(0) + _put_varint(len(key)) + _put_varint(len(value))) + len(key) + len(value)
    body += b"".join(struct.pack("<I", o) for o in (offs or [0]))
    body += struct.pack("<I", max(len(offs), 1))
    return bytes(body)


def write(prefix: str, tensors: Dict[str, np.ndarray], block_entries: int = 4):
    """Write ``tensors`` as a one-shard bundle (test fixture / export of the engine's weights for the reference)."""
    data = bytearray()
    items: List[Tuple[bytes, bytes]] = [(b"", b"\x08\x01\x1a\x02\x08\x01")]     # num_shards = 1, version.producer = 1
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        a = a if a.ndim == 0 or a.flags.c_contiguous else np.ascontiguousarray(a)
        if a.dtype not in DTYPE_IDS:
            raise RuntimeError("IGNNITION: dtype %s cannot be written to a checkpoint" % a.dtype)
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        items.append((name.encode(), _entry_proto(DTYPE_IDS[a.dtype], a.shape, len(data), len(raw))))
        data += raw
    out = bytearray()
    index_items = []
    for i in range(0, len(items), block_entries):
        blk = _build_block(items[i:i + block_entries])
        handle = _put_varint(len(out)) + _put_varint(len(blk))
        out += blk + b"\x00" + b"\x00\x00\x00\x00"            # no compression, crc not verified by the reader
        index_items.append((items[min(i + block_entries, len(items)) - 1][0], handle))
    meta = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out += meta + b"\x00" + b"\x00\x00\x00\x00"
    idx = _build_block(index_items)
    idx_handle = _put_varint(len(out)) + _put_varint(len(idx))
    out += idx + b"\x00" + b"\x00\x00\x00\x00"
    footer = meta_handle + idx_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
