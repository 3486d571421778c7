"""Reader / writer for TensorFlow "tensor bundle" checkpoints (``tf.train.Saver`` write_version V2:
``<prefix>.index`` + ``<prefix>.data-00000-of-00001``) without TensorFlow, so that models trained by
the reference (cfl/bin/train.py -> ``saver.save``, cfl/models/cfl.py:1430-1460, cfl/utils.py:465-497)
can be scored here and models trained here can be handed back.

Format, restated from TensorFlow's public sources (tensorflow/core/util/tensor_bundle, core/lib/io/
table*, which follow LevelDB's table format):
  * ``.index`` is an immutable sorted table: data blocks + meta-index block + index block + a 48-byte
    footer ``[metaindex handle][index handle][padding to 40 bytes][magic 0xdb4775248b80fb57]``; a
    handle is two varint64 (offset, size); every block is followed by a 1-byte compression type
    (0 = none; the bundle writer does not compress) and a 4-byte masked CRC32C; a block is a run of
    prefix-compressed entries ``varint shared | varint non_shared | varint value_len | key tail |
    value`` followed by the restart offsets (uint32 each) and their count (uint32).
  * key ``""`` -> ``BundleHeaderProto`` (num_shards = 1, endianness = 0 little, version.producer = 1);
    key ``<variable name>`` -> ``BundleEntryProto`` {1: dtype, 2: shape {2: dim {1: size}}, 3: shard_id,
    4: offset, 5: size, 6: fixed32 masked crc32c of the bytes}.
  * ``.data-00000-of-00001`` holds the raw little-endian tensor bytes at those offsets.
TensorFlow is not installable in the build container, so no TF-written file has been read.  What the
tests establish instead (tests/test_tf_checkpoint.py): known-answer CRC32C / varint vectors; the masked
CRC against the implementation in tensorboard's TensorFlow stub; the bundle protos in both directions
against Google's protobuf runtime over TensorFlow's own compiled DataType / TensorShapeProto / VersionDef
descriptors (shipped with tensorboard); round trips, multi-block tables and corruption detection for the
table layer, which has no independent implementation in this image.
"""
from __future__ import annotations

import os
import struct
from typing import Dict

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 10: np.dtype("bool")}
_DTYPE_IDS = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9,
              np.dtype("bool"): 10}


# ---- CRC32C (Castagnoli), masked as LevelDB / TF store it ------------------------------------------
def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_table()
_CRC_NP = np.array(_CRC_TABLE, dtype=np.uint32)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf ------------------------------------------------------------------------
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf: bytes, pos: int):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _pb_fields(buf: bytes):
    """(field number, wire type, value) of one protobuf message level."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _pb_varint(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field: int, b: bytes) -> bytes:
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _parse_entry(value: bytes):
    dtype, shape, shard, offset, size, crc = 0, [], 0, 0, 0, None
    for f, wt, v in _pb_fields(value):
        if f == 1:
            dtype = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:                                   # TensorShapeProto.dim
                    sz = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            sz = v3
                    shape.append(sz)
        elif f == 3:
            shard = v
        elif f == 4:
            offset = v
        elif f == 5:
            size = v
        elif f == 6:
            crc = struct.unpack("<I", v)[0]
        elif f == 7:
            raise NotImplementedError("partitioned (sliced) variables are not supported")
    return dtype, shape, shard, offset, size, crc


# ---- table reading ----------------------------------------------------------------------------------------
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    block, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if trailer[0] != 0:
        raise NotImplementedError("compressed table block (type %d): the TF bundle writer does not compress; "
                                  "this file was not written by tf.train.Saver" % trailer[0])
    if verify and struct.unpack("<I", trailer[1:5])[0] != masked_crc32c(block + trailer[:1]):
        raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
    return block


def _block_entries(block: bytes):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(path: str, verify: bool = True) -> Dict[str, tuple]:
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = buf[-48:]
    _, p = _get_varint(footer, 0)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    entries = {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, hp = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, hp)
        for key, value in _block_entries(_read_block(buf, boff, bsize, verify)):
            entries[key.decode("utf-8")] = value
    header = entries.pop("", None)
    if header is None:
        raise ValueError("%s: bundle header missing" % path)
    for f, _, v in _pb_fields(header):
        if f == 1 and v != 1:
            raise NotImplementedError("checkpoint sharded over %d data files" % v)
        if f == 2 and v != 0:
            raise NotImplementedError("big-endian checkpoint")
    return {k: _parse_entry(v) for k, v in entries.items()}


def load_tf_checkpoint(prefix: str, verify: bool = False) -> Dict[str, np.ndarray]:
    """All tensors of ``<prefix>.index`` / ``.data-00000-of-00001`` by variable name."""
    index = read_index(prefix + ".index", verify=True)
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    out = {}
    for name, (dtype, shape, shard, offset, size, crc) in index.items():
        if dtype not in _DTYPES:
            continue                                          # string / resource entries (e.g. _CHECKPOINTABLE_OBJECT_GRAPH)
        raw = data[offset:offset + size]
        if verify and crc is not None and masked_crc32c(raw) != crc:
            raise ValueError("checkpoint data: checksum mismatch for %s" % name)
        out[name] = np.frombuffer(raw, dtype=_DTYPES[dtype]).reshape(shape).copy()
    return out


# ---- writing ------------------------------------------------------------------------------------------------
def _build_block(items, restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_tf_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write ``tensors`` (by variable name) as a single-shard V2 checkpoint."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    header = _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1))
    items = [(b"", header)]
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        arr = np.asarray(tensors[name], order="C")                # (ascontiguousarray would turn scalars into [1])
        if arr.dtype not in _DTYPE_IDS:
            raise TypeError("%s: dtype %s not supported" % (name, arr.dtype))
        raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        shape = b"".join(_pb_bytes(2, _pb_varint(1, int(s))) for s in arr.shape)
        entry = _pb_varint(1, _DTYPE_IDS[arr.dtype]) + _pb_bytes(2, shape) + _pb_varint(4, len(data)) + \
            _pb_varint(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw))
        items.append((name.encode("utf-8"), entry))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    table = bytearray()

    def emit(block: bytes):
        off = len(table)
        table.extend(block + b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    data_handle = emit(_build_block(items))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block([(items[-1][0] + b"\x00", data_handle)], restart_interval=1))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    table.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(table))


# ---- model <-> checkpoint -------------------------------------------------------------------------------------
def adam_step_from_powers(tensors: Dict[str, np.ndarray], beta1: float = 0.9, beta2: float = 0.999):
    """Number of Adam updates t behind a TF checkpoint, from the float32 accumulators tf.train.AdamOptimizer
    stores (``beta_power = beta^(t+1)`` after t updates).  ``beta2_power`` is tried first: 0.999^t stays a
    normal float32 up to t ~ 87 000, whereas 0.9^t underflows to exactly 0 near t ~ 1000, so ``beta1_power``
    alone cannot date a realistically trained model.  Returns None when neither accumulator is usable."""
    import math
    for suffix, beta in (("beta2_power", beta2), ("beta1_power", beta1)):
        if not 0.0 < beta < 1.0:
            continue
        for k, v in tensors.items():
            if k.split("/")[-1] == suffix or k.endswith(suffix):
                x = float(np.asarray(v).reshape(-1)[0])
                if 1e-37 < x < 1.0:                               # normal float32 range: log is accurate to ~1e-7 relative
                    return max(0, int(round(math.log(x) / math.log(beta))) - 1)
    return None


def tf_to_state_dict(tensors: Dict[str, np.ndarray], beta1: float = 0.9, beta2: float = 0.999):
    """TF variable names are the ones our models register (tests/test_reference_golden.py); Adam's
    slot variables keep their TF names (``<var>/Adam``, ``<var>/Adam_1``); the step comes from the
    ``beta*_power`` accumulators (see ``adam_step_from_powers``).  When it cannot be recovered the slots
    are still loaded but a warning says that the bias correction restarts at step 0."""
    import warnings
    import torch
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in tensors.items()
          if not k.endswith("_power") and not k.endswith("_power_1") and "ExponentialMovingAverage" not in k}
    step = adam_step_from_powers(tensors, beta1, beta2)
    if step is not None:
        sd["__step__"] = torch.tensor(step)
    elif any(k.endswith("/Adam") for k in tensors):
        warnings.warn("TF checkpoint: Adam step not recoverable from beta1_power / beta2_power (underflowed or absent); "
                      "the moments are restored but the bias correction restarts at step 0")
    return sd


def state_dict_to_tf(sd, beta1: float = 0.9, beta2: float = 0.999):
    out = {}
    step = int(sd.get("__step__", 0))
    for k, v in sd.items():
        if k == "__step__":
            continue
        out[k] = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
    out["beta1_power"] = np.float32(beta1 ** (step + 1))
    out["beta2_power"] = np.float32(beta2 ** (step + 1))
    return out
