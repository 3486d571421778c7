"""cfl.input_data -- the reference's on-disk dataset layer for vector features
(cfl/input_data.py:195-265, 344-600; formats in SURVEY App. E), rebuilt for a GPU-resident catalog.

The reference opens ``features.b`` and seeks once PER ITEM PER BATCH (input_data.py:212-228) -- its
real wall-clock sink.  Here the whole file is mapped once (fixed stride 10 + 4F bytes: 10 ASCII bytes
of item id, then F little-endian float32), copied to the device once, and every batch is an index
gather on the GPU.  Batch composition (epoch wrap + reshuffle, oversampling when the batch is larger
than the pair list, ``data_switch`` swap) follows the reference's ``numpy.random.RandomState`` call
order exactly, so the same seed yields the same batches.  Image / "double" (PNG + latent) records
are out of scope (GAN / image half).
"""
from __future__ import annotations

import os
from argparse import Namespace

import numpy as np
import torch
from numpy.random import RandomState

ID_BYTES = 10


def _record_dtype(input_size):
    return np.dtype([("id", "S%d" % ID_BYTES), ("f", "<f4", (input_size,))])


def map_features(path, input_size):
    """Whole features.b as a structured memmap (ids, float rows)."""
    rec = _record_dtype(input_size)
    size = os.path.getsize(path)
    if size % rec.itemsize != 0:
        raise ValueError(f"{path}: size {size} is not a multiple of the record stride {rec.itemsize} "
                         f"(10-byte id + {input_size} float32)")
    return np.memmap(path, dtype=rec, mode="r")


def write_features(path, ids, features):
    """Writer for the same format (what convert_mnist.py:97-112 / monomer.patch:99-149 emit)."""
    features = np.asarray(features, dtype="<f4")
    rec = np.empty(len(ids), dtype=_record_dtype(features.shape[1]))
    for i, s in enumerate(ids):
        b = s.encode("ascii")
        if len(b) != ID_BYTES:
            raise ValueError(f"item id {s!r} must be exactly {ID_BYTES} ASCII characters")
        rec["id"][i] = b
    rec["f"] = features
    rec.tofile(path)


def load_features(path, input_size=28 * 28):
    """cfl/input_data.py:195-210: yields (asin, feature)."""
    for r in map_features(path, input_size):
        yield r["id"].decode("ascii"), np.array(r["f"])


def load_features_by_positions(path, positions, input_size=28 * 28):
    """cfl/input_data.py:212-228."""
    return np.array(map_features(path, input_size)["f"][np.asarray(positions, dtype=np.int64)])


def load_asins_by_positions(path, positions, input_size=28 * 28):
    """cfl/input_data.py:231-245."""
    ids = map_features(path, input_size)["id"]
    return [ids[int(p)].decode("ascii") for p in positions]


def load_features_indices(path, input_size=28 * 28):
    """cfl/input_data.py:248-265: asin -> position."""
    ids = map_features(path, input_size)["id"]
    return {a.decode("ascii"): i for i, a in enumerate(ids)}


def load_meta_lines(path):
    """cfl/input_data.py:268-287."""
    with open(path) as infile:
        lines, current_id = [], None
        for line in infile:
            if not line.startswith(" "):
                if current_id:
                    yield current_id, lines
                current_id = line.split(" ", 1)[0].strip()
                lines = [line]
            else:
                assert lines and current_id, "must have valid id"
                lines.append(line)
        if current_id:
            yield current_id, lines


def _read_pairs(path, asins_to_index):
    out = []
    with open(path) as infile:
        for line in infile:
            a, _, b = line.strip().split()                    # "<id_src> <relation> <id_tgt>"
            out.append([asins_to_index[a], asins_to_index[b]])
    return np.array(out, dtype=np.int64).reshape(-1, 2)


class SemiDataSet(object):
    """cfl/input_data.py:344-600 for vector features; batches are device tensors."""

    def __init__(self, path, input_size=28 * 28, data_switch=False, is_image=False, is_double=False,
                 directed=False, reorder=False, raw_latent=False, seed=633, device=None):
        if is_image or is_double:
            raise NotImplementedError("image / double records belong to the generation half (out of scope)")
        if reorder:
            raise NotImplementedError("reorder is a visualisation option of the generation half")
        self._rng = RandomState(seed)
        self.input_size = input_size
        self.feature_path = os.path.join(path, "features.b")
        self.is_image, self.is_double = False, False
        self.directed, self.data_switch, self.raw_latent = directed, data_switch, raw_latent
        mm = map_features(self.feature_path, input_size)
        ids = [a.decode("ascii") for a in mm["id"]]
        self.asins_to_index = {a: i for i, a in enumerate(ids)}
        self.index_to_asins = {i: a for a, i in self.asins_to_index.items()}
        # NOTE num_examples = max(index), not +1 (input_data.py:399-400): the last item never appears
        # in unlabeled batches.  Kept for batch-for-batch parity.
        self.num_examples = max(self.index_to_asins)
        self.item_indices = np.arange(self.num_examples)
        self.head_unlabeled = 0
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        host = torch.from_numpy(np.ascontiguousarray(mm["f"]))          # one sequential read of the file
        if self.device.type == "cuda":
            host = host.pin_memory()
        self.features = host.to(self.device, non_blocking=True)         # [n_items, F] resident
        if self.directed:
            def read_ids(name):
                with open(os.path.join(path, name)) as infile:
                    return np.array(sorted(self.asins_to_index[line.strip()] for line in infile))
            self.source_indices = read_ids("source.txt")
            self.num_source = self.source_indices.shape[0]
            self.source_indices = self.source_indices[self._rng.permutation(self.num_source)]
            self.head_source = 0
            self.target_indices = read_ids("target.txt")
            self.num_target = self.target_indices.shape[0]
            self.target_indices = self.target_indices[self._rng.permutation(self.num_target)]
            self.head_target = 0
        self.pairs_pos = _read_pairs(os.path.join(path, "pairs_pos.txt"), self.asins_to_index)
        self.pairs_neg = _read_pairs(os.path.join(path, "pairs_neg.txt"), self.asins_to_index)
        self.head_labeled_pos = 0
        self.head_labeled_neg = 0
        self.num_examples_labeled_pos = self.pairs_pos.shape[0]
        self.num_examples_labeled_neg = self.pairs_neg.shape[0]

    # -- gathers -------------------------------------------------------------------------------
    def _load_features_by_positions(self, indices):
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=self.device)
        return self.features.index_select(0, idx)

    def _load_asins_by_positions(self, indices):
        return [self.index_to_asins[int(i)] for i in indices]

    def next_batch(self, batch_size, return_labels=False):
        return self.next_labeled_batch(batch_size, return_labels)

    def next_labeled_batch(self, batch_size, return_labels=False):
        """(src_pos, dst_pos, src_neg, dst_neg), input_data.py:542-589."""
        if self.head_labeled_pos + batch_size > self.num_examples_labeled_pos:
            self.head_labeled_pos = 0
            self.pairs_pos = self.pairs_pos[self._rng.permutation(self.num_examples_labeled_pos)]
        if self.head_labeled_neg + batch_size > self.num_examples_labeled_neg:
            self.head_labeled_neg = 0
            self.pairs_neg = self.pairs_neg[self._rng.permutation(self.num_examples_labeled_neg)]
        positions_pos = self.pairs_pos[self.head_labeled_pos:self.head_labeled_pos + batch_size]
        positions_neg = self.pairs_neg[self.head_labeled_neg:self.head_labeled_neg + batch_size]
        if batch_size > self.num_examples_labeled_pos:
            positions_pos = self.pairs_pos[self._rng.choice(self.num_examples_labeled_pos, batch_size)]
        if batch_size > self.num_examples_labeled_neg:
            positions_neg = self.pairs_neg[self._rng.choice(self.num_examples_labeled_neg, batch_size)]
        assert positions_pos.shape[0] == batch_size and positions_neg.shape[0] == batch_size
        src_pos = self._load_features_by_positions(positions_pos[:, 0])
        dst_pos = self._load_features_by_positions(positions_pos[:, 1])
        src_neg = self._load_features_by_positions(positions_neg[:, 0])
        dst_neg = self._load_features_by_positions(positions_neg[:, 1])
        if self.data_switch and self._rng.rand() > 0.5:
            src_pos, dst_pos = dst_pos, src_pos
            src_neg, dst_neg = dst_neg, src_neg
        self.head_labeled_pos += batch_size
        self.head_labeled_neg += batch_size
        if return_labels:
            raise NotImplementedError()
        return src_pos, dst_pos, src_neg, dst_neg

    def next_unlabeled_batch(self, batch_size, return_labels=False, source_ids=False):
        """input_data.py:591-618."""
        if self.head_unlabeled + batch_size > self.num_examples:
            self.head_unlabeled = 0
            self.item_indices = self.item_indices[self._rng.permutation(self.num_examples)]
        positions = self.item_indices[self.head_unlabeled:self.head_unlabeled + batch_size]
        self.head_unlabeled += batch_size
        data = [self._load_features_by_positions(positions)]
        if source_ids:
            data.append(self._load_asins_by_positions(positions))
        return data

    def whole_unlabeled_batches(self, batch_size, source_ids=False):
        for i in range(0, self.num_examples, batch_size):
            positions = self.item_indices[i:i + batch_size]
            data = [self._load_features_by_positions(positions)]
            if source_ids:
                data.append(self._load_asins_by_positions(positions))
            yield data

    def whole_pos_batches(self, batch_size, source_ids=False):
        """input_data.py:503-520."""
        for i in range(0, self.num_examples_labeled_pos, batch_size):
            p = self.pairs_pos[i:i + batch_size]
            out = (self._load_features_by_positions(p[:, 0]), self._load_features_by_positions(p[:, 1]))
            yield out + (self._load_asins_by_positions(p[:, 0]),) if source_ids else out

    def whole_neg_batches(self, batch_size, source_ids=False):
        """input_data.py:522-540."""
        for i in range(0, self.num_examples_labeled_neg, batch_size):
            p = self.pairs_neg[i:i + batch_size]
            out = (self._load_features_by_positions(p[:, 0]), self._load_features_by_positions(p[:, 1]))
            yield out + (self._load_asins_by_positions(p[:, 0]),) if source_ids else out


def load_data_sets(path, input_size, data_switch=False, raw_latent=False, is_image=False, is_double=False,
                   directed=False, reorder=False, seed=633, device=None):
    """cfl/input_data.py:290-341: train / val / test splits."""
    mk = lambda split, **kw: SemiDataSet(path=os.path.join(path, split), input_size=input_size, is_image=is_image,
                                         is_double=is_double, raw_latent=raw_latent, directed=directed, seed=seed,
                                         device=device, **kw)
    return Namespace(train=mk("train", data_switch=data_switch), val=mk("val"), test=mk("test", reorder=reorder))
