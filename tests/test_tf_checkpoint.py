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
