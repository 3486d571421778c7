"""Monomer mode on the query x catalog cross product (cfl_score_topk_monomer, cfl.ranking.
MonomerCatalogIndex): the monomer branch of DistBase.build_dist (cfl/models/base.py:109-117) on
every (source = query, target = candidate) pair.

CPU part: the oracle's all-pairs restatement is the pair scorer on the cross product, and its
diagonal reproduces the distances the REFERENCE's own graph code produced for the monomer fixtures
(tests/golden/ref_cfl_monomer_*.npz).  GPU part: the CUDA kernel against that oracle through the C ABI."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cfl_oracle as O
from oracle import torch_port as T
from test_reference_golden import GOLDEN, encoder_params, load_case

MONO_CASES = sorted(glob.glob(os.path.join(GOLDEN, "ref_cfl_monomer_*.npz")))
RTOL = 1e-4                       # north star: distances within 1e-4 relative in fp32


def _case_operands(path, src_key="in_pos_source", tgt_key="in_pos_target"):
    """Query side (a, w) of the fixture's source batch and catalog side P of its target batch (fp64)."""
    z, cfg = load_case(path)
    names = ("DistEncoderSrc", "DistEncoderDst") if cfg["directed"] else ("DistEncoder", "DistEncoder")
    ps, pt = encoder_params(z, cfg, names[0]), encoder_params(z, cfg, names[1])
    scale = 1.0 / cfg["data_norm"][0] if cfg["data_norm"] else 1.0
    S = O.build_prototypes(z[src_key] * scale, ps, "monomer", cfg["K"], cfg["d"], cfg["act_type"])
    Tt = O.build_prototypes(z[tgt_key] * scale, pt, "monomer", cfg["K"], cfg["d"], cfg["act_type"])
    return z, cfg, ps, pt, scale, S["activations"], S["monomer_activations"], Tt["prototype_activations"]


# ------------------------------------------------------------------------------------ CPU
def test_all_pairs_monomer_is_the_pair_scorer_on_the_cross_product():
    rng = np.random.default_rng(23)
    Q, K, d, N = 6, 3, 5, 37
    a, Pt = rng.normal(size=(Q, d)), rng.normal(size=(N, K, d))
    w = O.softmax(rng.normal(size=(Q, K)))
    D = O.all_pairs_monomer_dist(a, w, Pt, block=4)
    for q in range(Q):
        want = O.monomer_dist(np.repeat(a[q:q + 1], N, 0), Pt, np.repeat(w[q:q + 1], N, 0))
        np.testing.assert_allclose(D[q], want, rtol=1e-13)
        tp = T.monomer_dist(torch.as_tensor(np.repeat(a[q:q + 1], N, 0)), torch.as_tensor(Pt),
                            torch.as_tensor(np.repeat(w[q:q + 1], N, 0)))
        np.testing.assert_allclose(D[q], tp.numpy().reshape(-1), rtol=1e-12)


def test_fixture_inventory_has_monomer_cases():
    assert len(MONO_CASES) >= 3


@pytest.mark.parametrize("path", MONO_CASES, ids=[os.path.basename(p)[8:-4] for p in MONO_CASES])
def test_cross_product_diagonal_is_the_reference_graph_distance(path):
    """Pair (i, i) of the cross product is the pair the reference's graph scored: the oracle's
    all-pairs restatement must return the reference's own s_pos_dists / s_neg_dists there."""
    for lab in ("pos", "neg"):
        z, cfg, *_, a, w, Pt = _case_operands(path, "in_%s_source" % lab, "in_%s_target" % lab)
        D = O.all_pairs_monomer_dist(a, w, Pt)
        np.testing.assert_allclose(np.diag(D), z["out_s_%s_dists" % lab][:, 0], rtol=1e-11, atol=1e-13)


# ------------------------------------------------------------------------------------ GPU
def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def host(t):
    return t.detach().cpu().numpy().astype(np.float64)


@pytest.fixture(scope="module")
def nat():
    from cfl import _native
    _native.lib()
    _, major, _ = _native.device_info()
    assert major == 10, "these tests need a B200 (sm_100)"
    return _native


def _inputs(rng, Q, K, d, N):
    Pt = rng.normal(size=(N, K, d)).astype(np.float32)
    a = (Pt[rng.integers(0, N, Q), rng.integers(0, K, Q)] + 0.5 * rng.normal(size=(Q, d))).astype(np.float32)
    w = O.softmax(2.0 * rng.normal(size=(Q, K))).astype(np.float32)
    return a, w, Pt


def _check(nat, a, w, Pt, k, idx_base=0):
    Q, N = a.shape[0], Pt.shape[0]
    D = O.all_pairs_monomer_dist(a.astype(np.float64), w.astype(np.float64), Pt.astype(np.float64), block=8)
    tv, ti, dense = nat.score_topk_monomer(dev(a), dev(w), dev(Pt), k, idx_base=idx_base, want_dense=True)
    tv, ti, dense = host(tv), ti.cpu().numpy(), host(dense)
    # (1) every score of the cross product, direct-difference form in fp32: 1e-4 relative (measured ~1e-6)
    err = np.abs(dense - D)
    assert (err <= RTOL * D + 1e-30).all(), f"max rel err {np.max(err / np.maximum(D, 1e-30)):.3e}"
    # (2) the reported values ARE the dense values of the reported rows (no second arithmetic)
    kk = min(k, N)
    loc = ti[:, :kk] - idx_base
    assert (loc >= 0).all() and (loc < N).all()
    assert (np.take_along_axis(dense, loc, 1) == tv[:, :kk]).all()
    # (3) the ranking of the kernel's own fp32 values is exact: ascending, ties -> lower index
    want_v32, want_i32 = O.rank_topk(dense, kk)
    assert (loc == want_i32).all() and (tv[:, :kk] == want_v32).all()
    # (4) against the fp64 ranking: index sets equal except candidates tied with the k-th at fp32 resolution
    want_v, want_i = O.rank_topk(D, kk)
    for q in range(Q):
        kth = want_v[q, -1]
        for c_ in set(loc[q].tolist()) ^ set(want_i[q].tolist()):
            assert abs(D[q, c_] - kth) <= 2e-6 * max(kth, 1.0) + 1e-7, f"q={q} cand {c_} is not a near-tie"
    if kk < k:
        assert (ti[:, kk:] == -1).all() and np.isinf(tv[:, kk:]).all()
    return tv, ti


@pytest.mark.gpu
@pytest.mark.parametrize("Q,K,d,N,k", [(1, 1, 1, 5, 3), (7, 2, 10, 1000, 100), (33, 4, 20, 5000, 100),
                                       (16, 4, 10, 20000, 100), (5, 8, 128, 1500, 50), (20, 3, 12, 129, 128),
                                       (3, 4, 15, 50, 100), (130, 3, 64, 2500, 20), (17, 5, 7, 128, 1)])
def test_score_topk_monomer_shapes(nat, Q, K, d, N, k):
    rng = np.random.default_rng(Q + K + d + N)
    _check(nat, *_inputs(rng, Q, K, d, N), k, idx_base=1000 if Q == 7 else 0)


@pytest.mark.gpu
def test_score_topk_monomer_long_parts_compact_their_buffers(nat):
    """Many tiles per catalog part: the running threshold / warp compaction path (cnt > 384)."""
    rng = np.random.default_rng(41)
    Q, K, d = 2400, 2, 8                    # 150 query tiles -> 3 catalog parts of ~52 tiles each
    a, w, Pt = _inputs(rng, Q, K, d, 20000)
    tv, ti, dense = nat.score_topk_monomer(dev(a), dev(w), dev(Pt), 100, want_dense=True)
    want_v, want_i = O.rank_topk(host(dense), 100)
    assert (ti.cpu().numpy() == want_i).all() and (host(tv) == want_v).all()
    sub = slice(0, 2400, 97)
    D = O.all_pairs_monomer_dist(a[sub].astype(np.float64), w[sub].astype(np.float64), Pt.astype(np.float64), block=2)
    assert (np.abs(host(dense)[sub] - D) <= RTOL * D).all()


@pytest.mark.gpu
def test_score_topk_monomer_exact_ties_and_determinism(nat):
    rng = np.random.default_rng(42)
    a, w, Pt = _inputs(rng, 12, 3, 16, 700)
    Pt[7] = Pt[7, 1][None, :]                   # all K prototypes of row 7 coincide: tiny distance to a[:6]
    Pt[100] = Pt[7]
    Pt[650] = Pt[7]
    a[:6] = Pt[7, 1] + 0.01 * rng.normal(size=(6, 16)).astype(np.float32)      # rows 7/100/650 rank high
    tv, ti = _check(nat, a, w, Pt, 20)
    seen = 0
    for q in range(12):
        pos = {c: int(np.where(ti[q] == c)[0][0]) for c in (7, 100, 650) if c in ti[q]}
        if len(pos) == 3:
            seen += 1
            assert pos[7] + 1 == pos[100] and pos[100] + 1 == pos[650]
    assert seen >= 6
    b = nat.score_topk_monomer(dev(a), dev(w), dev(Pt), 20)
    assert (host(b[0]) == tv).all() and (b[1].cpu().numpy() == ti).all()


@pytest.mark.gpu
def test_score_topk_monomer_matches_the_paired_kernel(nat):
    """The value reported for (q, c) is the paired monomer kernel's distance of that pair (1e-6)."""
    rng = np.random.default_rng(43)
    Q, K, d, N, k = 40, 4, 20, 3000, 10
    a, w, Pt = _inputs(rng, Q, K, d, N)
    tv, ti = nat.score_topk_monomer(dev(a), dev(w), dev(Pt), k)
    rows = ti.reshape(-1)
    aa = dev(a).repeat_interleave(k, 0)
    ww = dev(w).repeat_interleave(k, 0)
    dist, _, _, _ = nat.pair_loss_fwd("monomer", aa, dev(Pt)[rows], w=ww)
    np.testing.assert_allclose(host(dist), host(tv).reshape(-1), rtol=5e-6)


@pytest.mark.gpu
def test_score_topk_monomer_shards_merge_to_the_single_list(nat):
    rng = np.random.default_rng(44)
    Q, K, d, N, k, R = 24, 4, 20, 8000, 100, 4
    a, w, Pt = (dev(x) for x in _inputs(rng, Q, K, d, N))
    full_v, full_i = nat.score_topk_monomer(a, w, Pt, k)
    vs, is_ = [], []
    for r in range(R):
        lo, hi = r * N // R, (r + 1) * N // R
        v, i = nat.score_topk_monomer(a, w, Pt[lo:hi], k, idx_base=lo)
        vs.append(v); is_.append(i)
    mv, mi = nat.topk_merge(torch.stack(vs), torch.stack(is_))
    assert torch.equal(mi, full_i) and torch.equal(mv, full_v)


@pytest.mark.gpu
def test_score_topk_monomer_empty_and_ragged(nat):
    a, w, Pt = (dev(x) for x in _inputs(np.random.default_rng(45), 5, 2, 6, 9))
    tv, ti = nat.score_topk_monomer(a[:0], w[:0], Pt, 4)
    assert tv.shape == (0, 4) and ti.shape == (0, 4)
    tv, ti = nat.score_topk_monomer(a, w, Pt[:0], 4)
    assert (ti == -1).all() and torch.isinf(tv).all()
    # strided views: query rows and catalog rows with a leading dimension larger than the row
    big_a = torch.zeros(5, 11, device="cuda"); big_a[:, :6] = a
    big_P = torch.zeros(9, 20, device="cuda"); big_P[:, :12] = Pt.reshape(9, 12)
    v1, i1 = nat.score_topk_monomer(a, w, Pt, 4)
    v2, i2 = nat.score_topk_monomer(big_a[:, :6], w, big_P[:, :12].unflatten(1, (2, 6)), 4)
    assert torch.equal(v1, v2) and torch.equal(i1, i2)
    with pytest.raises(nat.CflNativeError):
        nat.score_topk_monomer(a, w[:, :1], Pt, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("path", MONO_CASES, ids=[os.path.basename(p)[8:-4] for p in MONO_CASES])
def test_monomer_index_reproduces_the_reference_graph_distances(nat, path):
    """cfl.ranking.MonomerCatalogIndex on the fixture's weights: catalog = the target batch, queries = the
    source batch; entry (i, i) of the cross product is the reference graph's own distance of pair i."""
    from cfl.ranking import EncoderWeights, MonomerCatalogIndex
    z, cfg, ps, pt, scale, a64, w64, Pt64 = _case_operands(path)
    V0, g0, b0 = ps["outputs"]
    Vg, gg, _ = ps["monomer_outputs"]
    Vp, gp, bp = pt["prototype_outputs"]
    opt = lambda v: None if v is None else dev(v)
    wts = EncoderWeights(V0=dev(V0), Vp=dev(Vp), g0=dev(g0), gp=dev(gp), b0=opt(b0), bp=opt(bp), weight_norm=True,
                         in_scale=scale, act=cfg["act_type"], Vg=dev(Vg), gg=dev(gg))
    index = MonomerCatalogIndex.from_features(wts, dev(z["in_pos_target"]))
    n = z["in_pos_target"].shape[0]
    aq, gate = index.project_queries(dev(z["in_pos_source"]))
    np.testing.assert_allclose(host(aq), a64, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(host(gate), w64, rtol=1e-4, atol=1e-6)
    tv, ti = index.rank(dev(z["in_pos_source"]), k=n)
    tv, ti = host(tv), ti.cpu().numpy()
    ref = z["out_s_pos_dists"][:, 0]
    for i in range(n):
        j = int(np.where(ti[i] == i)[0][0])
        np.testing.assert_allclose(tv[i, j], ref[i], rtol=RTOL)
    D = O.all_pairs_monomer_dist(a64, w64, Pt64)
    np.testing.assert_allclose(tv, np.take_along_axis(D, ti, 1), rtol=RTOL)


@pytest.mark.gpu
def test_monomer_full_size_catalog_properties(nat):
    """BASELINE config 2 shape (Monomer: K=4, d=20) at a 1M-item catalog: size-independent properties."""
    g = torch.Generator(device="cuda").manual_seed(633)
    N, K, d, Q, k = 1_000_000, 4, 20, 256, 100
    Pt = torch.randn(N, K, d, generator=g, device="cuda")
    a = Pt[torch.randint(0, N, (Q,), generator=g, device="cuda"), 0] + 0.5 * torch.randn(Q, d, generator=g, device="cuda")
    w = torch.softmax(2 * torch.randn(Q, K, generator=g, device="cuda"), -1)
    tv, ti = nat.score_topk_monomer(a, w, Pt, k)
    assert (tv[:, 1:] >= tv[:, :-1]).all()                                    # sorted
    assert (ti >= 0).all() and (ti < N).all()
    assert all(len(set(r.tolist())) == k for r in ti[:16].cpu())              # no duplicates
    # reported values = the paired kernel on the gathered rows
    rows = ti.reshape(-1)
    dist, _, _, _ = nat.pair_loss_fwd("monomer", a.repeat_interleave(k, 0), Pt[rows], w=w.repeat_interleave(k, 0))
    torch.testing.assert_close(dist, tv.reshape(-1), rtol=5e-6, atol=0)
    # sharding the catalog and merging gives the same list, bit for bit
    vs, is_ = [], []
    for r in range(8):
        lo, hi = r * N // 8, (r + 1) * N // 8
        v, i = nat.score_topk_monomer(a, w, Pt[lo:hi], k, idx_base=lo)
        vs.append(v); is_.append(i)
    mv, mi = nat.topk_merge(torch.stack(vs), torch.stack(is_))
    assert torch.equal(mi, ti) and torch.equal(mv, tv)
    # nothing outside the list beats the k-th: a fp32 torch evaluation of 4 queries over the whole catalog
    for q in range(4):
        dq = (w[q][None, :] * ((a[q][None, None, :] - Pt) ** 2).sum(-1)).sum(-1)
        kth = torch.topk(dq, k, largest=False).values[-1]
        assert abs(float(kth) - float(tv[q, -1])) <= 1e-5 * float(kth)


@pytest.mark.gpu
def test_score_topk_monomer_golden_fixture(nat):
    """The committed fixture (tests/golden/rank_monomer.npz, frozen oracle output incl. exact duplicate rows)."""
    g = np.load(os.path.join(GOLDEN, "rank_monomer.npz"))
    k = int(g["k"])
    tv, ti = _check(nat, g["a"], g["w"], g["Pt"], k)
    assert (ti == g["top_idx"]).mean() >= 0.99            # _check already bounds every mismatch to an fp32 near-tie
    np.testing.assert_allclose(tv, g["top_val"], rtol=RTOL, atol=1e-7)
    for q in range(4):                          # rows 7 / 90 / 555 are identical: ties -> lower index
        assert ti[q, :3].tolist() == [7, 90, 555]


# ---------------------------------------------------------------------------- tensor-core path
@pytest.mark.gpu
@pytest.mark.parametrize("Q,K,d,N,k,clustered", [(70, 4, 20, 150_000, 100, False), (130, 2, 10, 60_000, 100, False),
                                                  (33, 4, 20, 40_000, 17, True), (9, 1, 64, 30_000, 100, False),
                                                  (50, 6, 20, 50_000, 100, False), (5, 4, 20, 900, 100, False)])
def test_score_topk_monomer_packed_equals_the_exact_kernel(nat, Q, K, d, N, k, clustered):
    """Tensor-core route (Gram filter over the augmented vectors + exact rescoring + verification / redo) == the
    CUDA-core kernel, bit for bit (values are rescored in its arithmetic), and both match the fp64 oracle."""
    rng = np.random.default_rng(1000 + Q + K)
    a, w, Pt = _inputs(rng, Q, K, d, N)
    if clustered:                              # tight clusters: thousands of rows within a hair of every threshold
        cen = rng.normal(size=(20, K, d)).astype(np.float32)
        Pt = (cen[rng.integers(0, 20, N)] + 0.02 * rng.normal(size=(N, K, d))).astype(np.float32)
        a = (cen[rng.integers(0, 20, Q), 0] + 0.05 * rng.normal(size=(Q, d))).astype(np.float32)
    Pd = dev(Pt)
    mu = Pd.reshape(N, K, d).mean(dim=(0, 1))
    img = nat.monomer_pack(Pd, mu)
    assert img is not None
    ev, ei = nat.score_topk_monomer(dev(a), dev(w), Pd, k)
    tv, ti, st = nat.score_topk_monomer_packed(dev(a), dev(w), Pd, img, k, mu=mu, want_stats=True)
    assert torch.equal(ei, ti) and torch.equal(ev, tv)
    stats = dict(zip(nat.SCORE_STAT_NAMES, st.tolist()))
    assert stats["lower_bound_pass"] == 1
    if not clustered and N >= 30_000:
        assert stats["redo_queries"] <= max(1, Q // 20), stats
    qs = [0, Q // 2, Q - 1]
    D = O.all_pairs_monomer_dist(a[qs].astype(np.float64), w[qs].astype(np.float64), Pt.astype(np.float64), block=1)
    wv, wi = O.rank_topk(D, min(k, N))
    np.testing.assert_allclose(host(tv[qs])[:, :wv.shape[1]], wv, rtol=RTOL)


@pytest.mark.gpu
def test_score_topk_monomer_packed_range_guard_and_large_offset(nat):
    """A prototype value outside the fp16 range (flag -> every query redone exactly) and an uncentred catalog far
    from the origin (mu = None: large margins) still return the exact kernel's lists."""
    rng = np.random.default_rng(77)
    Q, K, d, N, k = 40, 4, 20, 50_000, 50
    a, w, Pt = _inputs(rng, Q, K, d, N)
    Pt = Pt + 10.0
    a = a + 10.0
    Pd = dev(Pt)
    img = nat.monomer_pack(Pd, None)
    ev, ei = nat.score_topk_monomer(dev(a), dev(w), Pd, k)
    tv, ti = nat.score_topk_monomer_packed(dev(a), dev(w), Pd, img, k, mu=None)
    assert torch.equal(ei, ti) and torch.equal(ev, tv)
    Pd[123, 2, 5] = 3.0e5
    img = nat.monomer_pack(Pd, None)
    ev, ei = nat.score_topk_monomer(dev(a), dev(w), Pd, k)
    tv, ti, st = nat.score_topk_monomer_packed(dev(a), dev(w), Pd, img, k, mu=None, want_stats=True)
    assert torch.equal(ei, ti) and torch.equal(ev, tv)
    assert dict(zip(nat.SCORE_STAT_NAMES, st.tolist()))["redo_queries"] == Q


@pytest.mark.gpu
def test_monomer_index_uses_the_tensor_core_path_and_agrees(nat):
    from cfl.ranking import EncoderWeights, MonomerCatalogIndex
    rng = np.random.default_rng(5)
    F, K, d, N, Q = 64, 4, 20, 40_000, 48
    w = EncoderWeights(V0=dev(O.xavier_uniform(rng, F, d)), Vp=dev(O.xavier_uniform(rng, F, K * d)), g0=torch.ones(d).cuda(),
                       gp=torch.ones(K * d).cuda(), weight_norm=True, Vg=dev(O.xavier_uniform(rng, d, K)), gg=torch.ones(K).cuda())
    X = dev(rng.normal(size=(N, F)).astype(np.float32))
    idx = MonomerCatalogIndex.from_features(w, X)
    assert idx.image is not None
    tv, ti = idx.rank(X[:Q], 100)
    a, gate = idx.project_queries(X[:Q])
    ev, ei = nat.score_topk_monomer(a, gate, idx.P, 100)
    assert torch.equal(ti, ei) and torch.equal(tv, ev)
