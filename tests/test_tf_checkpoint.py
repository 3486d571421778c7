"""TensorFlow V2 checkpoint (tensor bundle) reader / writer: known-answer vectors of the primitives the
format is built from, structural checks of the files, round trips, corruption detection, and the
Saver / load_model path.  No TensorFlow here: see the module docstring for what this
does and does not establish."""
import os
import struct

import numpy as np
import pytest
import torch


def test_crc32c_and_varint_known_answers():
    from cfl import tf_checkpoint as T
    assert T.crc32c(b"123456789") == 0xE3069283                      # the standard CRC-32C check value
    assert T.crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4 test patterns
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    c = T.crc32c(b"foo")
    assert T.masked_crc32c(b"foo") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    for v, enc in ((0, b"\x00"), (127, b"\x7f"), (128, b"\x80\x01"), (300, b"\xac\x02"), (2 ** 32, b"\x80\x80\x80\x80\x10")):
        assert T._put_varint(v) == enc and T._get_varint(enc, 0) == (v, len(enc))


def _tensors(rng, n_extra=40):
    t = {"CFL/DistEncoder/outputs/fully_connected/V": rng.normal(size=(24, 8)).astype(np.float32),
         "CFL/DistEncoder/outputs/fully_connected/g": rng.normal(size=8).astype(np.float32),
         "CFL/DistEncoder/outputs/fully_connected/biases": np.zeros(8, np.float32),
         "CFL/Thresholder/threshold/threshold": np.float32(0.37),
         "CFL/beta1_power": np.float32(0.9 ** 8), "CFL/beta2_power": np.float32(0.999 ** 8),
         "global_step": np.int64(7)}
    for i in range(n_extra):                                          # several restart intervals, shared prefixes
        t["CFL/DistEncoder/prototype_outputs/fully_connected/V/Adam_%d" % i] = rng.normal(size=(3, i % 5 + 1)).astype(np.float32)
    return t


def test_round_trip_structure_and_corruption(tmp_path):
    from cfl import tf_checkpoint as T
    rng = np.random.default_rng(0)
    tensors = _tensors(rng)
    prefix = str(tmp_path / "model-120")
    T.write_tf_checkpoint(prefix, tensors)
    assert sorted(os.listdir(tmp_path)) == ["model-120.data-00000-of-00001", "model-120.index"]
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw[-48:]) == 48
    back = T.load_tf_checkpoint(prefix, verify=True)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(back[k], v)
    idx = T.read_index(prefix + ".index")
    assert list(idx) == sorted(idx, key=lambda s: s.encode())          # table keys are sorted
    off = idx["CFL/Thresholder/threshold/threshold"][3]
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    assert struct.unpack("<f", data[off:off + 4])[0] == np.float32(0.37)
    data[off] ^= 0xFF
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        T.load_tf_checkpoint(prefix, verify=True)
    bad = bytearray(raw); bad[10] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        T.read_index(prefix + ".index")
    open(prefix + ".index", "wb").write(raw[:-1] + b"\x00")
    with pytest.raises(ValueError, match="magic"):
        T.read_index(prefix + ".index")


class _Model:
    beta1, beta2 = 0.9, 0.999

    def __init__(self):
        self.sd = {}

    def state_dict(self):
        return self.sd

    def load_state_dict(self, sd):
        self.sd = dict(sd)


def test_saver_restores_reference_style_checkpoints(tmp_path):
    """load_model picks up `<dir>/checkpoint` -> `model-N.index` exactly like a directory the reference wrote."""
    from cfl import tf_checkpoint as T
    from cfl.utils import Session, export_tf_checkpoint, load_model
    rng = np.random.default_rng(1)
    m = _Model()
    m.sd = {"CFL/DistEncoder/outputs/fully_connected/V": torch.tensor(rng.normal(size=(6, 4)), dtype=torch.float32),
            "CFL/DistEncoder/outputs/fully_connected/V/Adam": torch.ones(6, 4),
            "CFL/DistEncoder/outputs/fully_connected/V/Adam_1": torch.full((6, 4), 2.0),
            "CFL/Thresholder/threshold/threshold": torch.tensor(0.5), "__step__": torch.tensor(41)}
    ck = tmp_path / "ck"
    export_tf_checkpoint(m, str(ck / "model-599"))
    tf_vars = T.load_tf_checkpoint(str(ck / "model-599"))
    assert float(tf_vars["beta1_power"]) == pytest.approx(0.9 ** 42, rel=1e-6)    # TF stores beta^(t+1)
    m2 = _Model()
    _, start = load_model(Session(m2), str(ck))
    assert start == 600
    assert int(m2.sd["__step__"]) == 41
    for k in m.sd:
        if k != "__step__":
            assert torch.equal(m2.sd[k], m.sd[k]), k


def test_adam_step_survives_beta1_power_underflow():
    """float32 0.9^t is exactly 0 from t ~ 1000 on; the step must then come from beta2_power."""
    import warnings
    import numpy as np
    from cfl import tf_checkpoint as T
    for t in (0, 7, 999, 5000, 60000):
        tensors = {"beta1_power": np.float32(0.9) ** np.float32(t + 1), "beta2_power": np.float32(0.999 ** (t + 1)),
                   "Dist/Encoder/latent_outputs/fully_connected/weights": np.zeros((2, 2), np.float32),
                   "Dist/Encoder/latent_outputs/fully_connected/weights/Adam": np.zeros((2, 2), np.float32)}
        sd = T.tf_to_state_dict(tensors)
        got = int(sd["__step__"])
        assert abs(got - t) <= max(1, int(2e-4 * t)), (t, got)   # float32 log resolution
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sd = T.tf_to_state_dict({"beta1_power": np.float32(0.0), "x/Adam": np.zeros(2, np.float32)})
        assert "__step__" not in sd and any("not recoverable" in str(x.message) for x in w)


def _bundle_proto_classes():
    """BundleHeaderProto / BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto) as real protobuf messages:
    the two message layouts are declared here, but their sub-messages and enums -- DataType, TensorShapeProto,
    VersionDef -- are TensorFlow's own compiled descriptors as shipped in the tensorboard package, and encoding /
    decoding is done by Google's protobuf runtime, not by cfl.tf_checkpoint's hand-written varint code."""
    pytest.importorskip("google.protobuf")
    pytest.importorskip("tensorboard")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2
    pool = descriptor_pool.Default()
    try:
        pool.FindMessageTypeByName("tensorboard.BundleEntryProto")
    except KeyError:
        fd = descriptor_pb2.FileDescriptorProto(name="cfl_tests/tensor_bundle.proto", package="tensorboard", syntax="proto3")
        fd.dependency.extend([tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
        hdr = fd.message_type.add(name="BundleHeaderProto")
        en = hdr.enum_type.add(name="Endianness")
        en.value.add(name="LITTLE", number=0)
        en.value.add(name="BIG", number=1)
        ent = fd.message_type.add(name="BundleEntryProto")
        F = descriptor_pb2.FieldDescriptorProto
        for msg, fields in ((hdr, (("num_shards", 1, F.TYPE_INT32, None),
                                   ("endianness", 2, F.TYPE_ENUM, ".tensorboard.BundleHeaderProto.Endianness"),
                                   ("version", 3, F.TYPE_MESSAGE, ".tensorboard.VersionDef"))),
                            (ent, (("dtype", 1, F.TYPE_ENUM, ".tensorboard.DataType"),
                                   ("shape", 2, F.TYPE_MESSAGE, ".tensorboard.TensorShapeProto"),
                                   ("shard_id", 3, F.TYPE_INT32, None), ("offset", 4, F.TYPE_INT64, None),
                                   ("size", 5, F.TYPE_INT64, None), ("crc32c", 6, F.TYPE_FIXED32, None)))):
            for name, number, typ, type_name in fields:
                f = msg.field.add(name=name, number=number, type=typ, label=F.LABEL_OPTIONAL)
                if type_name:
                    f.type_name = type_name
        pool.Add(fd)
    get = message_factory.GetMessageClass
    return (get(pool.FindMessageTypeByName("tensorboard.BundleHeaderProto")),
            get(pool.FindMessageTypeByName("tensorboard.BundleEntryProto")), types_pb2)


def _raw_index_entries(T, path):
    """key -> raw value bytes of an .index table (the walk of read_index without the protobuf parsing)."""
    buf = open(path, "rb").read()
    footer = buf[-48:]
    _, p = T._get_varint(footer, 0)
    _, p = T._get_varint(footer, p)
    ioff, p = T._get_varint(footer, p)
    isize, p = T._get_varint(footer, p)
    out = {}
    for _, handle in T._block_entries(T._read_block(buf, ioff, isize, True)):
        boff, hp = T._get_varint(handle, 0)
        bsize, _ = T._get_varint(handle, hp)
        for key, value in T._block_entries(T._read_block(buf, boff, bsize, True)):
            out[key.decode()] = value
    return out


def test_crc_and_bundle_protos_against_independent_implementations(tmp_path):
    """Two of the format's three layers against implementations that are not ours (TensorFlow itself is absent):
    the masked CRC-32C against the one in tensorboard's TensorFlow stub (written by the TF team for reading TF event
    files), and the bundle protos against Google's protobuf runtime over TensorFlow's compiled DataType /
    TensorShapeProto / VersionDef descriptors -- in both directions: what write_tf_checkpoint emits parses to the
    intended messages, and entries serialised by the protobuf runtime the way a proto3 writer does (zero fields
    omitted, empty shape message for scalars) are read back by load_tf_checkpoint.  The table layer (LevelDB block
    format) stays validated by structure checks and round trips only."""
    from cfl import tf_checkpoint as T
    Header, Entry, types_pb2 = _bundle_proto_classes()
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb
    rng = np.random.default_rng(11)
    for n in (0, 1, 3, 4, 5, 31, 32, 33, 1000, 4097):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert T.masked_crc32c(data) == tb.masked_crc32c(data), n
        assert T.crc32c(data) == tb.crc32c(data) & 0xFFFFFFFF, n

    # (1) our writer -> protobuf runtime
    tensors = _tensors(rng)
    tensors["flags/is_double"] = np.array([True, False])
    tensors["stats/f64"] = rng.normal(size=(2, 3))
    tensors["stats/i32"] = np.arange(5, dtype=np.int32)
    prefix = str(tmp_path / "model-7")
    T.write_tf_checkpoint(prefix, tensors)
    raw = _raw_index_entries(T, prefix + ".index")
    hdr = Header.FromString(raw.pop(""))
    assert hdr.num_shards == 1 and hdr.endianness == 0 and hdr.version.producer == 1
    blob = open(prefix + ".data-00000-of-00001", "rb").read()
    names = {np.dtype("float32"): "DT_FLOAT", np.dtype("float64"): "DT_DOUBLE", np.dtype("int32"): "DT_INT32",
             np.dtype("int64"): "DT_INT64", np.dtype("bool"): "DT_BOOL"}
    assert set(raw) == set(tensors)
    covered = 0
    for name, value in raw.items():
        e = Entry.FromString(value)
        want = np.asarray(tensors[name])
        assert types_pb2.DataType.Name(e.dtype) == names[want.dtype], name
        assert [d.size for d in e.shape.dim] == list(want.shape) and not e.shape.unknown_rank, name
        assert e.shard_id == 0 and e.size == want.nbytes, name
        chunk = blob[e.offset:e.offset + e.size]
        assert chunk == want.tobytes() and e.crc32c == tb.masked_crc32c(chunk), name
        assert Entry.FromString(value).SerializeToString() == Entry.FromString(e.SerializeToString()).SerializeToString()
        covered += e.size
    assert covered == len(blob)                                     # the data file is exactly the tensors, no gaps

    # (2) protobuf runtime -> our reader: entries as a proto3 writer serialises them, several data blocks
    prefix2 = str(tmp_path / "model-8")
    dt = {v: getattr(types_pb2, v) for v in names.values()}
    items, data = [(b"", Header(num_shards=1, endianness=0, version={"producer": 1}).SerializeToString())], bytearray()
    for name in sorted(tensors, key=lambda s: s.encode()):
        arr = np.asarray(tensors[name])
        e = Entry(dtype=dt[names[arr.dtype]], offset=len(data), size=arr.nbytes, crc32c=tb.masked_crc32c(arr.tobytes()))
        e.shape.SetInParent()                                       # TF writes the shape message even for scalars
        for s in arr.shape:
            e.shape.dim.add().size = int(s)
        items.append((name.encode(), e.SerializeToString()))
        data += arr.tobytes()
    first = [f for f, _, _ in T._pb_fields(items[1][1])]
    assert 4 not in first and 3 not in first and 2 in first          # offset 0 / shard 0 are omitted on the wire
    table, handles = bytearray(), []

    def emit(block):
        off = len(table)
        table.extend(block + b"\x00" + struct.pack("<I", tb.masked_crc32c(block + b"\x00")))
        return T._put_varint(off) + T._put_varint(len(block))

    for lo in range(0, len(items), 7):                              # 7 entries per data block
        part = items[lo:lo + 7]
        handles.append((part[-1][0] + b"\x00", emit(T._build_block(part))))
    meta = emit(T._build_block([]))
    index = emit(T._build_block(handles, restart_interval=1))
    footer = meta + index
    table.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.TABLE_MAGIC))
    open(prefix2 + ".index", "wb").write(bytes(table))
    open(prefix2 + ".data-00000-of-00001", "wb").write(bytes(data))
    back = T.load_tf_checkpoint(prefix2, verify=True)
    assert set(back) == set(tensors)
    for name, want in tensors.items():
        want = np.asarray(want)
        assert back[name].dtype == want.dtype and back[name].shape == want.shape and np.array_equal(back[name], want), name
