"""Per-query all-candidate AUC (cfl_pair_dist_rows + cfl_rank_counts + cfl.ranking.auc_from_rank_counts):
the rank statistic of roc_auc_score (cfl/utils.py:267-268) for one query against every catalog row.

CPU: the integer bookkeeping against the oracle's exact AUC and sklearn (the function the reference calls),
and the sharded path on a world-size-2 gloo group with oracle-backed kernel doubles.  GPU: the kernels
through the C ABI against the fp64 oracle, bit-exact on integer-valued inputs (massive ties)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cfl_oracle as O


def _counts_from_dense(D, pos_idx):
    """[Q,J,2] (lt, eq) over all candidates and the thresholds, from a dense distance matrix (numpy)."""
    Q, J = pos_idx.shape
    t = np.full((Q, J), np.nan, dtype=D.dtype)
    cnt = np.zeros((Q, J, 2), dtype=np.int64)
    for q in range(Q):
        for j in range(J):
            if pos_idx[q, j] >= 0:
                t[q, j] = D[q, pos_idx[q, j]]
                cnt[q, j, 0] = (D[q] < t[q, j]).sum()
                cnt[q, j, 1] = (D[q] == t[q, j]).sum()
    return cnt, t


def _planted(rng, Q, N, J, pad=True):
    pos = np.stack([rng.choice(N, J, replace=False) for _ in range(Q)]).astype(np.int64)
    if pad:
        pos[::3, -1] = -1                       # ragged: some queries have J-1 positives
        if Q > 4:
            pos[4, :] = -1                      # a query without positives
    return pos


# ------------------------------------------------------------------------------------ CPU
def test_auc_from_rank_counts_equals_exact_auc_and_sklearn():
    from sklearn.metrics import roc_auc_score
    from cfl.ranking import auc_from_rank_counts
    rng = np.random.default_rng(3)
    Q, N, J = 9, 500, 6
    D = np.round(rng.gamma(2.0, size=(Q, N)), 1).astype(np.float32)       # many exact ties
    pos = _planted(rng, Q, N, J)
    D[0, pos[0, :3]] = D[0, pos[0, 0]]                                     # ties among the positives themselves
    cnt, t = _counts_from_dense(D, pos)
    auc, two_u, n_pos, n_neg = auc_from_rank_counts(torch.as_tensor(cnt), torch.as_tensor(t), N)
    wu, wp, wn = O.per_query_auc(D.astype(np.float64), pos)
    assert two_u.tolist() == wu and n_pos.tolist() == wp and n_neg.tolist() == wn
    for q in range(Q):
        if wp[q] == 0:
            assert np.isnan(float(auc[q]))
            continue
        labels = np.zeros(N, dtype=int)
        labels[pos[q][pos[q] >= 0]] = 1
        assert abs(float(auc[q]) - roc_auc_score(labels, -D[q].astype(np.float64))) <= 4 * np.finfo(np.float64).eps


def _doubles(nat):
    def pair_dist_rows(mode, query, catalog, pos_idx, w=None):
        D = O.all_pairs_dist(query.numpy().astype(np.float64), catalog.numpy().astype(np.float64)).astype(np.float32)
        n = catalog.shape[0]
        out = np.full(pos_idx.shape, np.nan, dtype=np.float32)
        for q in range(pos_idx.shape[0]):
            for j in range(pos_idx.shape[1]):
                r = int(pos_idx[q, j])
                if 0 <= r < n:
                    out[q, j] = D[q, r]
        return torch.as_tensor(out)

    def rank_counts(mode, query, catalog, pos_dist, w=None):
        D = O.all_pairs_dist(query.numpy().astype(np.float64), catalog.numpy().astype(np.float64)).astype(np.float32)
        t = pos_dist.numpy()
        with np.errstate(invalid="ignore"):
            lt = (D[:, None, :] < t[:, :, None]).sum(-1)
            eq = (D[:, None, :] == t[:, :, None]).sum(-1)
        return torch.as_tensor(np.stack([lt, eq], -1).astype(np.int64))

    nat.pair_dist_rows, nat.rank_counts = pair_dist_rows, rank_counts


def _case():
    rng = np.random.default_rng(11)
    N, K, d, Q, J = 300, 3, 6, 7, 4
    E = np.round(rng.normal(size=(N, d)), 1).astype(np.float32)
    Pq = np.round(rng.normal(size=(Q, K, d)), 1).astype(np.float32)
    return E, Pq, _planted(rng, Q, N, J)


def _auc_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "compatibility-family-learning_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cfl import ranking
    _doubles(ranking.nat)
    E, Pq, pos = _case()
    lo, hi = ranking.shard_bounds(E.shape[0], world, rank)
    r = ranking._auc_per_query("pcd", torch.as_tensor(Pq), torch.as_tensor(E[lo:hi]), None, torch.as_tensor(pos),
                               lo, E.shape[0], None, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "auc.npz"), auc=r.auc.numpy(), two_u=r.two_u.numpy(), counts=r.counts.numpy(),
                 pos_dist=r.pos_dist.numpy())
    dist.destroy_process_group()


def test_sharded_auc_world2_gloo_matches_single_process(tmp_path):
    """Positives' distances are taken where they live, counts add over shards: same integers as one process."""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_auc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "auc.npz")
    E, Pq, pos = _case()
    D = O.all_pairs_dist(Pq.astype(np.float64), E.astype(np.float64)).astype(np.float32)
    cnt, t = _counts_from_dense(D, pos)
    assert (got["counts"] == cnt).all()
    np.testing.assert_array_equal(got["pos_dist"], t)
    wu, wp, wn = O.per_query_auc(D.astype(np.float64), pos)
    assert got["two_u"].tolist() == wu


# ------------------------------------------------------------------------------------ GPU
def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


@pytest.fixture(scope="module")
def nat():
    from cfl import _native
    _native.lib()
    _, major, _ = _native.device_info()
    assert major == 10, "these tests need a B200 (sm_100)"
    return _native


def _bracket(D, t, rel):
    """Counts every fp32 evaluation within `rel` of the fp64 distances must satisfy: lt in [lo_lt, hi_le]."""
    with np.errstate(invalid="ignore"):
        lo = (D[:, None, :] < (t * (1 - rel))[:, :, None]).sum(-1)
        hi = (D[:, None, :] <= (t * (1 + rel))[:, :, None]).sum(-1)
    return lo, hi


def _check_counts(nat, mode, query, catalog, gate, D, pos, rel=2e-5):
    pos_t = torch.as_tensor(pos).cuda()
    g = None if gate is None else dev(gate)
    t = nat.pair_dist_rows(mode, dev(query), dev(catalog), pos_t, w=g)
    cnt = nat.rank_counts(mode, dev(query), dev(catalog), t, w=g).cpu().numpy()
    t = t.cpu().numpy().astype(np.float64)
    valid = pos >= 0
    assert (np.isnan(t) == ~valid).all()
    want_t = np.where(valid, np.take_along_axis(D, np.maximum(pos, 0), 1), np.nan)
    np.testing.assert_allclose(t[valid], want_t[valid], rtol=1e-4)                      # the fp32 bar of the north star
    assert (cnt[~valid] == 0).all()
    assert (cnt[..., 1][valid] >= 1).all(), "a positive must compare equal to itself"
    lo, hi = _bracket(D, want_t, rel)
    assert (cnt[..., 0][valid] >= lo[valid]).all() and ((cnt[..., 0] + cnt[..., 1])[valid] <= hi[valid]).all()
    return t, cnt


@pytest.mark.gpu
@pytest.mark.parametrize("Q,K,d,N,J", [(1, 1, 1, 5, 1), (9, 3, 64, 3000, 8), (33, 4, 20, 5000, 8), (5, 8, 128, 700, 3),
                                       (20, 2, 12, 129, 32), (130, 3, 16, 2500, 4)])
def test_rank_counts_pcd_against_the_oracle(nat, Q, K, d, N, J):
    rng = np.random.default_rng(Q + K + d + N)
    E = rng.normal(size=(N, d)).astype(np.float32)
    Pq = (E[rng.integers(0, N, Q)][:, None, :] + 0.5 * rng.normal(size=(Q, K, d))).astype(np.float32)
    pos = _planted(rng, Q, N, min(J, N), pad=N > 5)
    D = O.all_pairs_dist(Pq.astype(np.float64), E.astype(np.float64))
    _check_counts(nat, "pcd", Pq, E, None, D, pos)


@pytest.mark.gpu
@pytest.mark.parametrize("Q,K,d,N,J", [(9, 4, 20, 3000, 8), (17, 3, 8, 1000, 5), (4, 8, 32, 600, 2)])
def test_rank_counts_monomer_against_the_oracle(nat, Q, K, d, N, J):
    rng = np.random.default_rng(Q + K + d + N + 1)
    Pt = rng.normal(size=(N, K, d)).astype(np.float32)
    a = (Pt[rng.integers(0, N, Q), 0] + 0.5 * rng.normal(size=(Q, d))).astype(np.float32)
    w = O.softmax(rng.normal(size=(Q, K))).astype(np.float32)
    pos = _planted(rng, Q, N, J)
    D = O.all_pairs_monomer_dist(a.astype(np.float64), w.astype(np.float64), Pt.astype(np.float64))
    _check_counts(nat, "monomer", a, Pt, w, D, pos)
    # same arithmetic as the ranking kernel: the distances of ranked rows are the ranking's values, bit for bit,
    # and a ranked row's strict count is its position in the list (no ties in this random data)
    tv, ti = nat.score_topk_monomer(dev(a), dev(w), dev(Pt), 10)
    t = nat.pair_dist_rows("monomer", dev(a), dev(Pt), ti, w=dev(w))
    assert torch.equal(t, tv)
    cnt = nat.rank_counts("monomer", dev(a), dev(Pt), t, w=dev(w))
    assert (cnt[..., 0].cpu() == torch.arange(10)[None, :]).all() and (cnt[..., 1] == 1).all()


@pytest.mark.gpu
def test_rank_counts_bit_exact_on_integer_inputs_with_massive_ties(nat):
    """Small-integer embeddings and coinciding prototypes: every fp32 distance is an exact integer, so the
    counts and the AUC integers must equal the oracle's exactly (siamese/K=1 and K=2,4 with s = 1/K)."""
    from cfl.ranking import auc_from_rank_counts
    rng = np.random.default_rng(7)
    N, d, Q, J = 4000, 10, 24, 6
    E = rng.integers(-2, 3, size=(N, d)).astype(np.float32)
    for K in (1, 2, 4):
        p = rng.integers(-2, 3, size=(Q, 1, d)).astype(np.float32)
        Pq = np.repeat(p, K, axis=1)
        pos = _planted(rng, Q, N, J)
        D = O.all_pairs_dist(Pq.astype(np.float64), E.astype(np.float64))
        assert (D == np.round(D)).all()
        want_cnt, want_t = _counts_from_dense(D, pos)
        t = nat.pair_dist_rows("pcd" if K > 1 else "siamese", dev(Pq), dev(E), torch.as_tensor(pos).cuda())
        cnt = nat.rank_counts("pcd" if K > 1 else "siamese", dev(Pq), dev(E), t)
        np.testing.assert_array_equal(t.cpu().numpy().astype(np.float64), want_t)
        assert (cnt.cpu().numpy() == want_cnt).all()
        auc, two_u, n_pos, n_neg = auc_from_rank_counts(cnt, t, N)
        wu, wp, wn = O.per_query_auc(D, pos)
        assert two_u.tolist() == wu and n_pos.tolist() == wp and n_neg.tolist() == wn


@pytest.mark.gpu
def test_rank_counts_add_over_catalog_shards(nat):
    rng = np.random.default_rng(8)
    N, K, d, Q, J, R = 9000, 3, 64, 40, 8, 4
    E = dev(rng.normal(size=(N, d)).astype(np.float32))
    Pq = dev(rng.normal(size=(Q, K, d)).astype(np.float32))
    pos = torch.as_tensor(_planted(rng, Q, N, J)).cuda()
    t = nat.pair_dist_rows("pcd", Pq, E, pos)
    full = nat.rank_counts("pcd", Pq, E, t)
    acc = torch.zeros_like(full)
    tt = torch.full_like(t, float("nan"))
    for r in range(R):
        lo, hi = r * N // R, (r + 1) * N // R
        loc = torch.where((pos >= lo) & (pos < hi), pos - lo, torch.full_like(pos, -1))
        part = nat.pair_dist_rows("pcd", Pq, E[lo:hi], loc)
        tt = torch.where(torch.isnan(part), tt, part)
        acc += nat.rank_counts("pcd", Pq, E[lo:hi], t)
    assert torch.equal(acc, full)
    assert torch.equal(torch.nan_to_num(tt, nan=-1.0), torch.nan_to_num(t, nan=-1.0))


@pytest.mark.gpu
def test_catalog_index_auc_per_query_planted_positives(nat):
    """SURVEY 8d C3 check at a reduced catalog: J=8 planted positives per query (near-duplicates of one of the
    query's prototypes in embedding space) get AUC ~ 1; values equal the oracle's AUC of the same distances."""
    from cfl.ranking import CatalogIndex, EncoderWeights
    rng = np.random.default_rng(633)
    F, K, d, N, Q, J = 64, 3, 16, 20000, 32, 8
    V0 = O.xavier_uniform(rng, F, d)
    Vp = O.xavier_uniform(rng, F, K * d)
    w = EncoderWeights(V0=dev(V0), Vp=dev(Vp), g0=torch.ones(d).cuda(), gp=torch.ones(K * d).cuda(),
                       b0=torch.zeros(d).cuda(), bp=torch.zeros(K * d).cuda())
    X = np.maximum(rng.normal(size=(N, F)), 0).astype(np.float32)
    index = CatalogIndex.from_features(w, dev(X))
    xq = X[:Q]
    Pq = index.project_queries(dev(xq))
    pos = rng.permutation(np.arange(Q, N))[:Q * J].reshape(Q, J).astype(np.int64)   # disjoint sets, not the queries' own rows
    # plant: overwrite the positives' embeddings with noisy copies of the query's prototypes
    E = index.E
    for q in range(Q):
        E[torch.as_tensor(pos[q]).cuda()] = Pq[q, rng.integers(0, K, J)] + 0.01 * torch.randn(J, d, device="cuda")
    index.image = nat.catalog_pack(E, K, index.mu)           # the embeddings changed: repack the catalog image
    r = index.auc_per_query(dev(xq), torch.as_tensor(pos))
    assert (r.auc > 0.999).all()
    assert torch.equal(r.counts, index.auc_per_query(dev(xq), torch.as_tensor(pos), method="direct").counts)
    D = O.all_pairs_dist(Pq.cpu().numpy().astype(np.float64), E.cpu().numpy().astype(np.float64))
    wu, wp, wn = O.per_query_auc(D, pos)
    want = np.array(wu) / (2.0 * np.array(wp) * np.array(wn))
    np.testing.assert_allclose(r.auc.cpu().numpy(), want, atol=2.0 / (J * (N - J)))   # at most a near-tie flip or two


@pytest.mark.gpu
@pytest.mark.parametrize("Q,N,J", [(1, 1, 1), (7, 1000, 8), (33, 4097, 9), (5, 70001, 20), (3, 513, 32)])
def test_dense_rank_counts_exact_on_tied_values(nat, Q, N, J):
    rng = np.random.default_rng(Q + N + J)
    D = np.round(rng.gamma(2.0, size=(Q, N)), 1).astype(np.float32)          # many exact ties
    pos = _planted(rng, Q, N, min(J, N), pad=N > 40)
    want_cnt, t = _counts_from_dense(D, pos)
    got = nat.dense_rank_counts(dev(D), dev(t))
    assert (got.cpu().numpy() == want_cnt).all()
    # a strided view (leading dimension > N) counts only the first N columns
    big = torch.full((Q, N + 13), -1.0, device="cuda")
    big[:, :N] = dev(D)
    assert torch.equal(nat.dense_rank_counts(big[:, :N], dev(t)), got)


@pytest.mark.gpu
def test_catalog_index_auc_per_query_fused_route(nat):
    """The tensor-core route (counts taken in the scoring kernel's epilogue, near-ties re-evaluated in the direct form)
    returns the SAME integers as the direct route; both agree with the fp64 oracle up to near-tie flips."""
    from cfl.ranking import CatalogIndex, EncoderWeights
    rng = np.random.default_rng(634)
    F, K, d, N, Q, J = 64, 3, 64, 30000, 48, 8
    w = EncoderWeights(V0=dev(O.xavier_uniform(rng, F, d)), Vp=dev(O.xavier_uniform(rng, F, K * d)),
                       g0=torch.ones(d).cuda(), gp=torch.ones(K * d).cuda(), b0=torch.zeros(d).cuda(),
                       bp=torch.zeros(K * d).cuda())
    X = np.maximum(rng.normal(size=(N, F)), 0).astype(np.float32)
    index = CatalogIndex.from_features(w, dev(X))
    xq = dev(X[:Q])
    pos = rng.permutation(np.arange(Q, N))[:Q * J].reshape(Q, J).astype(np.int64)
    pos[::5, -1] = -1
    a = index.auc_per_query(xq, torch.as_tensor(pos), method="direct")
    b = index.auc_per_query(xq, torch.as_tensor(pos))                      # default = fused
    assert torch.equal(a.counts, b.counts)
    assert torch.equal(torch.nan_to_num(a.pos_dist, nan=-1.0), torch.nan_to_num(b.pos_dist, nan=-1.0))
    assert (b.counts[..., 1][torch.as_tensor(pos >= 0).cuda()] >= 1).all()
    Pq = index.project_queries(xq)
    D = O.all_pairs_dist(Pq.cpu().numpy().astype(np.float64), index.E.cpu().numpy().astype(np.float64))
    wu, wp, wn = O.per_query_auc(D, pos)
    want = np.array(wu) / (2.0 * np.array(wp) * np.array(wn))
    np.testing.assert_allclose(b.auc.cpu().numpy(), want, atol=4.0 / (J * (N - J)))   # a few near-tie flips vs fp64


def _fused_case(rng, K, d, Q, N, J, kind):
    if kind == "integer":                       # integer-valued inputs: exact ties everywhere, every tie is ambiguous
        E = rng.integers(-2, 3, size=(N, d)).astype(np.float32)
        Pq = rng.integers(-2, 3, size=(Q, K, d)).astype(np.float32)
    else:
        E = rng.normal(size=(N, d)).astype(np.float32)
        Pq = (E[rng.integers(0, N, Q)][:, None, :] + 0.5 * rng.normal(size=(Q, K, d))).astype(np.float32)
        if kind == "offset":                    # a common offset eats mantissa unless the image is centred
            E += 10.0
            Pq += 10.0
        if kind == "far":                       # prototypes far apart: one-hot soft-min
            Pq *= 4.0
        if kind == "degenerate" and K > 1:      # coinciding prototypes: no affine-hull bound for those queries
            Pq[::2, 1] = Pq[::2, 0]
    pos = _planted(rng, Q, N, J, pad=Q > 4)
    if kind in ("top", "degenerate"):           # positives among the query's best rows: the group-skip regime
        D = O.all_pairs_dist(Pq.astype(np.float64), E.astype(np.float64))
        best = np.argsort(D, axis=1)[:, :4 * J]
        pos = np.stack([rng.permutation(best[q])[:J] for q in range(Q)]).astype(np.int64)
        pos[::3, -1] = -1
    return E, Pq, pos


@pytest.mark.gpu
@pytest.mark.parametrize("K,d,Q,N,J,kind", [
    (3, 64, 70, 30000, 8, "normal"), (1, 20, 5, 5000, 3, "normal"), (2, 16, 33, 20000, 9, "normal"),
    (4, 20, 64, 25000, 8, "offset"), (8, 20, 10, 9000, 8, "normal"), (5, 32, 50, 12000, 8, "normal"),
    (3, 128, 40, 10000, 20, "normal"), (4, 128, 9, 7000, 32, "far"), (2, 8, 20, 3000, 4, "integer"),
    (3, 12, 130, 129, 2, "integer"), (1, 64, 3, 1, 1, "normal"), (6, 10, 65, 40000, 8, "offset"),
    (3, 64, 70, 30000, 8, "top"), (4, 20, 40, 25000, 8, "top"), (1, 32, 33, 20000, 5, "top"),
    (3, 16, 50, 20000, 8, "degenerate")])
def test_rank_counts_packed_equals_direct(nat, K, d, Q, N, J, kind):
    """cfl_rank_counts_packed == cfl_rank_counts, integer for integer (ties, padded slots, ragged tiles, J > 8 in
    several threshold chunks), and the rounding band has headroom: no ambiguous pair deviated by half of it."""
    rng = np.random.default_rng(K * 1000 + d + Q + J)
    E, Pq, pos = _fused_case(rng, K, d, Q, N, min(J, N), kind)
    Ed, Pd = dev(E), dev(Pq)
    mu = nat.col_mean(Ed)
    img = nat.catalog_pack(Ed, K, mu)
    t = nat.pair_dist_rows("pcd", Pd, Ed, torch.as_tensor(pos).cuda())
    want = nat.rank_counts("pcd", Pd, Ed, t)
    got, st = nat.rank_counts_packed(Pd, Ed, img, mu, t, want_stats=True)
    assert torch.equal(got, want), (st, (got != want).sum().item())
    assert st["recounted_queries"] == 0
    if kind != "integer":
        assert st["worst_ratio"] < 0.5 and st["beyond_half_band"] == 0, st
    # counts at the fp64 oracle's precision: only near-ties may differ
    if N <= 30000 and Q <= 70:
        D = O.all_pairs_dist(Pq.astype(np.float64), E.astype(np.float64))
        cnt, _ = _counts_from_dense(D, pos)
        lt = got[..., 0].cpu().numpy()
        ok = pos >= 0
        assert np.abs(lt[ok] - cnt[..., 0][ok]).max() <= (N if kind == "integer" else 3)


@pytest.mark.gpu
def test_rank_counts_packed_overflowing_record_list_falls_back(nat, monkeypatch):
    """A record list that cannot hold the ambiguous pairs flags the affected queries; they are recounted from scratch
    on the CUDA cores and the integers still equal the direct route's."""
    monkeypatch.setenv("CFL_EXPERIMENTS", "1")
    monkeypatch.setenv("CFL_RANK_REC_CAP", "64")
    rng = np.random.default_rng(77)
    K, d, Q, N, J = 3, 16, 40, 20000, 8
    E, Pq, pos = _fused_case(rng, K, d, Q, N, J, "integer")
    Ed, Pd = dev(E), dev(Pq)
    mu = nat.col_mean(Ed)
    img = nat.catalog_pack(Ed, K, mu)
    t = nat.pair_dist_rows("pcd", Pd, Ed, torch.as_tensor(pos).cuda())
    want = nat.rank_counts("pcd", Pd, Ed, t)
    got, st = nat.rank_counts_packed(Pd, Ed, img, mu, t, want_stats=True)
    assert st["recounted_queries"] > 0
    assert torch.equal(got, want)


def test_auc_from_rank_counts_random_cases_against_the_exact_auc():
    """Randomised bookkeeping check (ties among positives, between positives and negatives, padded slots)."""
    from cfl.ranking import auc_from_rank_counts
    rng = np.random.default_rng(99)
    for trial in range(60):
        Q, N, J = int(rng.integers(1, 6)), int(rng.integers(3, 60)), int(rng.integers(1, 5))
        J = min(J, N - 1)
        D = rng.integers(0, 4, size=(Q, N)).astype(np.float32)             # four distinct values: ties everywhere
        pos = np.stack([rng.choice(N, J, replace=False) for _ in range(Q)]).astype(np.int64)
        pos[rng.uniform(size=pos.shape) < 0.2] = -1
        cnt, t = _counts_from_dense(D, pos)
        auc, two_u, n_pos, n_neg = auc_from_rank_counts(torch.as_tensor(cnt), torch.as_tensor(t), N)
        wu, wp, wn = O.per_query_auc(D.astype(np.float64), pos)
        assert two_u.tolist() == wu and n_pos.tolist() == wp and n_neg.tolist() == wn, trial
