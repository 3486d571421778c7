"""CPU tests of the dataset layer: byte-exact features.b format and batch-for-batch
agreement with an independent emulation of the reference's SemiDataSet batching rules."""
import os
import struct

import numpy as np
import pytest
from numpy.random import RandomState


def _make_split(d, n=23, F=6, n_pos=17, n_neg=31, seed=0, directed=False):
    from cfl import input_data as I
    rng = np.random.default_rng(seed)
    ids = ["B%09d" % i for i in range(n)]
    feats = rng.normal(size=(n, F)).astype(np.float32)
    os.makedirs(d, exist_ok=True)
    I.write_features(os.path.join(d, "features.b"), ids, feats)
    pos = rng.integers(0, n, size=(n_pos, 2))
    neg = rng.integers(0, n, size=(n_neg, 2))
    for name, pairs in (("pairs_pos.txt", pos), ("pairs_neg.txt", neg)):
        with open(os.path.join(d, name), "w") as f:
            for a, b in pairs:
                f.write(f"{ids[a]} match {ids[b]}\n")
    if directed:
        open(os.path.join(d, "source.txt"), "w").write("\n".join(ids[: n // 2]) + "\n")
        open(os.path.join(d, "target.txt"), "w").write("\n".join(ids[n // 2:]) + "\n")
    return ids, feats, pos, neg


def test_features_b_is_byte_exact(tmp_path):
    """10 ASCII bytes of id then F little-endian float32, fixed stride (input_data.py:195-228)."""
    from cfl import input_data as I
    ids, feats, _, _ = _make_split(str(tmp_path))
    raw = open(tmp_path / "features.b", "rb").read()
    stride = 10 + 4 * feats.shape[1]
    assert len(raw) == stride * len(ids)
    for p in (0, 7, len(ids) - 1):
        assert raw[p * stride:p * stride + 10].decode("ascii") == ids[p]
        vals = struct.unpack("<%df" % feats.shape[1], raw[p * stride + 10:(p + 1) * stride])
        np.testing.assert_array_equal(np.float32(vals), feats[p])
    np.testing.assert_array_equal(I.load_features_by_positions(tmp_path / "features.b", [3, 1], feats.shape[1]), feats[[3, 1]])
    assert I.load_asins_by_positions(tmp_path / "features.b", [5, 0], feats.shape[1]) == [ids[5], ids[0]]
    assert I.load_features_indices(tmp_path / "features.b", feats.shape[1])[ids[9]] == 9
    got = list(I.load_features(tmp_path / "features.b", feats.shape[1]))
    assert got[4][0] == ids[4] and np.array_equal(got[4][1], feats[4])
    with pytest.raises(ValueError):
        I.map_features(tmp_path / "features.b", feats.shape[1] + 1)


def _reference_batches(pos, neg, seed, batch_size, steps, data_switch):
    """Independent emulation of input_data.py:542-589 (same RandomState call order)."""
    rng = RandomState(seed)
    pos, neg = pos.copy(), neg.copy()
    hp = hn = 0
    out = []
    for _ in range(steps):
        if hp + batch_size > len(pos):
            hp = 0
            pos = pos[rng.permutation(len(pos))]
        if hn + batch_size > len(neg):
            hn = 0
            neg = neg[rng.permutation(len(neg))]
        pp, nn = pos[hp:hp + batch_size], neg[hn:hn + batch_size]
        if batch_size > len(pos):
            pp = pos[rng.choice(len(pos), batch_size)]
        if batch_size > len(neg):
            nn = neg[rng.choice(len(neg), batch_size)]
        swap = data_switch and rng.rand() > 0.5
        hp += batch_size
        hn += batch_size
        out.append((pp[:, 1], pp[:, 0], nn[:, 1], nn[:, 0]) if swap else (pp[:, 0], pp[:, 1], nn[:, 0], nn[:, 1]))
    return out


@pytest.mark.parametrize("batch_size,data_switch", [(5, False), (5, True), (17, True), (20, False), (40, True)])
def test_labeled_batches_follow_the_reference_rules(tmp_path, batch_size, data_switch):
    from cfl import input_data as I
    ids, feats, pos, neg = _make_split(str(tmp_path))
    ds = I.SemiDataSet(str(tmp_path), input_size=feats.shape[1], data_switch=data_switch, seed=633, device="cpu")
    want = _reference_batches(pos, neg, 633, batch_size, 12, data_switch)
    for w in want:
        got = ds.next_labeled_batch(batch_size)
        for g, idx in zip(got, w):
            np.testing.assert_array_equal(g.numpy(), feats[idx])


def test_whole_batches_unlabeled_and_directed(tmp_path):
    from cfl import input_data as I
    ids, feats, pos, neg = _make_split(str(tmp_path), directed=True)
    ds = I.SemiDataSet(str(tmp_path), input_size=feats.shape[1], directed=True, seed=1, device="cpu")
    got = list(ds.whole_pos_batches(6, source_ids=True))
    assert sum(len(b[0]) for b in got) == len(pos)
    np.testing.assert_array_equal(got[1][1].numpy(), feats[pos[6:12, 1]])
    assert got[0][2] == [ids[i] for i in pos[:6, 0]]
    assert sum(len(b[0]) for b in ds.whole_neg_batches(7)) == len(neg)
    # reference quirk kept: num_examples = max(index), the last item is never served unlabeled
    assert ds.num_examples == len(ids) - 1
    seen = np.concatenate([b[0].numpy() for b in ds.whole_unlabeled_batches(5)])
    assert len(seen) == len(ids) - 1
    assert sorted(ds.source_indices.tolist()) == list(range(len(ids) // 2)) and ds.num_target == len(ids) - len(ids) // 2
    x, = ds.next_unlabeled_batch(4)
    assert x.shape == (4, feats.shape[1])


def test_load_data_sets_and_eval_loop_shapes(tmp_path):
    from cfl import input_data as I
    for split in ("train", "val", "test"):
        _make_split(str(tmp_path / split), seed=hash(split) % 100)
    data = I.load_data_sets(str(tmp_path), 6, data_switch=True, seed=633, device="cpu")
    assert data.train.data_switch and not data.val.data_switch
    b = data.train.next_batch(4)
    assert len(b) == 4 and all(t.shape == (4, 6) for t in b)
    with pytest.raises(NotImplementedError):
        I.SemiDataSet(str(tmp_path / "train"), input_size=6, is_image=True)
