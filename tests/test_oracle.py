"""Pins the oracle (no GPU): closed forms vs an independent torch fp64 autograd
transcription of the reference call sites, vs sklearn's roc_auc_score (the function the
reference calls, cfl/utils.py:267-268), and vs the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import cfl_oracle as O
from oracle import torch_port as T

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RNG = lambda s=633: np.random.default_rng(s)   # 633 = reference default seed, cfl/utils.py:89


def _t(a, grad=False):
    x = torch.tensor(a, dtype=torch.float64)
    return x.requires_grad_(grad)


@pytest.mark.parametrize("K,d", [(1, 8), (2, 10), (3, 64), (4, 20), (5, 12), (8, 128)])
def test_pcd_dist_matches_torch_port_fwd_bwd(K, d):
    rng = RNG(K * 100 + d)
    B = 37
    v = rng.normal(size=(B, d))
    P = v[:, None, :] + 0.7 * rng.normal(size=(B, K, d))
    up = rng.normal(size=B)
    dist = O.pcd_dist(v, P)
    tv, tP = _t(v, True), _t(P, True)
    td = T.pcd_dist(tv, tP)
    np.testing.assert_allclose(dist, td.detach().numpy(), rtol=1e-12, atol=1e-12)
    (td * _t(up)).sum().backward()
    dv, dP = O.pcd_dist_bwd(v, P, up)
    np.testing.assert_allclose(dv, tv.grad.numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dP, tP.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_monomer_and_siamese_match_torch_port():
    rng = RNG(7)
    B, K, d = 29, 4, 20
    a, b = rng.normal(size=(B, d)), rng.normal(size=(B, d))
    Pt = rng.normal(size=(B, K, d))
    w = O.softmax(rng.normal(size=(B, K)))
    up = rng.normal(size=B)
    ta, tPt, tw = _t(a, True), _t(Pt, True), _t(w, True)
    td = T.monomer_dist(ta, tPt, tw)
    np.testing.assert_allclose(O.monomer_dist(a, Pt, w), td.detach().numpy(), rtol=1e-12)
    (td * _t(up)).sum().backward()
    da, dPt, dw = O.monomer_dist_bwd(a, Pt, w, up)
    np.testing.assert_allclose(da, ta.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dPt, tPt.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dw, tw.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(O.siamese_dist(a, b),
                               T.siamese_dist(_t(a), _t(b)).numpy(), rtol=1e-12)


@pytest.mark.parametrize("opts", [
    dict(),                                                   # Dist model (dist.py:253-269)
    dict(pos_weight=0.0625),                                  # dyadic/run.sh
    dict(pos_weight=0.25, lambda_m=0.5),
    dict(use_threshold=False, caffe_margin=2.0),
    dict(pos_weight=0.5, caffe_margin=1.5),
])
@pytest.mark.parametrize("theta", [1e-6, 0.5, -3.0, 2.0])
def test_losses_and_grads_match_torch_port(opts, theta):
    rng = RNG(11)
    dp = np.abs(rng.normal(size=50)) * 2
    dn = np.abs(rng.normal(size=50)) * 3
    L = O.dist_losses(dp, dn, np.float64(theta), reg=0.125, **opts)
    tdp, tdn, tth = _t(dp, True), _t(dn, True), _t(np.float64(theta), True)
    tot, lp, ln = T.dist_total_loss(tdp, tdn, tth, reg=0.125, **opts)
    np.testing.assert_allclose(L["total_loss"], tot.item(), rtol=1e-12)
    np.testing.assert_allclose(L["p_loss_pos"], lp.item(), rtol=1e-12)
    np.testing.assert_allclose(L["p_loss_neg"], ln.item(), rtol=1e-12)
    tot.backward()
    gp, gn, gth = O.dist_losses_bwd(dp, dn, np.float64(theta), **opts)
    np.testing.assert_allclose(gp, tdp.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(gn, tdn.grad.numpy(), rtol=1e-10, atol=1e-14)
    tg = 0.0 if tth.grad is None else tth.grad.item()
    np.testing.assert_allclose(gth, tg, rtol=1e-10, atol=1e-14)


def test_theta_tie_gets_gradient():
    """theta is initialised exactly at the clamp (blocks.py:20); TF's maximum routes the
    gradient to theta on the tie, so training can leave 1e-6."""
    dp, dn = np.array([0.3, 0.1]), np.array([0.9, 2.0])
    _, _, g = O.dist_losses_bwd(dp, dn, np.float64(1e-6))
    assert g != 0.0
    _, _, g0 = O.dist_losses_bwd(dp, dn, np.float64(1e-7))
    assert g0 == 0.0


def test_fc_weight_norm_fwd_bwd_matches_torch_port():
    rng = RNG(3)
    B, F, N = 23, 40, 12
    x = rng.normal(size=(B, F))
    V, g, b = rng.normal(size=(F, N)), rng.uniform(0.5, 2, N), rng.normal(size=N)
    dy = rng.normal(size=(B, N))
    y, pre, z = O.fc_weight_norm(x, V, g, b, None, return_pre=True)
    tV, tg, tb = _t(V, True), _t(g, True), _t(b, True)
    ty = T.fc_weight_norm(_t(x), tV, tg, tb)
    np.testing.assert_allclose(y, ty.detach().numpy(), rtol=1e-12)
    (ty * _t(dy)).sum().backward()
    dV, dg, db = O.fc_weight_norm_bwd(x, V, g, z, dy)
    np.testing.assert_allclose(dV, tV.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dg, tg.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(db, tb.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_build_prototypes_bias_rule_and_shapes():
    rng = RNG(5)
    B, F, K, d = 9, 30, 3, 8
    x = rng.normal(size=(B, F))
    mk = lambda n, bias: (rng.normal(size=(F, n)), np.ones(n), np.zeros(n) if bias else None)
    out = O.build_prototypes(x, {"outputs": mk(d, True), "prototype_outputs": mk(K * d, True)},
                             "pcd", K, d, act="tanh")
    assert out["prototype_activations"].shape == (B, K, d)
    # prototype k of row b lives at columns [k*d,(k+1)*d) (base.py:80-82)
    np.testing.assert_array_equal(out["prototype_activations"][:, 1, :],
                                  out["flat_prototype_activations"][:, d:2 * d])
    pm = {"outputs": mk(d, False), "prototype_outputs": mk(K * d, False),
          "monomer_outputs": (rng.normal(size=(d, K)), np.ones(K), None)}
    om = O.build_prototypes(x, pm, "monomer", K, d)
    np.testing.assert_allclose(om["monomer_activations"].sum(-1), 1.0, rtol=1e-12)


def test_identities_used_by_the_gram_kernel():
    """Gram expansion and translation invariance (fp64)."""
    rng = RNG(9)
    B, K, d = 64, 4, 20
    v, P = rng.normal(size=(B, d)) + 3, rng.normal(size=(B, K, d)) + 3
    dist, s, dk = O.pcd_dist(v, P, return_aux=True)
    G = np.einsum("bkd,bd->bk", P, v)
    pp = np.einsum("bkd,bld->bkl", P, P)
    gram = (v * v).sum(-1) - 2 * (s * G).sum(-1) + np.einsum("bk,bkl,bl->b", s, pp, s)
    np.testing.assert_allclose(gram, dist, rtol=1e-11)
    mu = rng.normal(size=d)
    np.testing.assert_allclose(O.pcd_dist(v - mu, P - mu), dist, rtol=1e-11)
    lower = dk.min(-1) - 0.5 * (1 - 1 / K) * np.max(
        ((P[:, :, None] - P[:, None]) ** 2).sum(-1), axis=(1, 2))
    assert (dist >= lower - 1e-9).all() and (dist <= (s * dk).sum(-1) + 1e-9).all()


def test_auc_exact_vs_sklearn():
    from sklearn.metrics import roc_auc_score
    rng = RNG(13)
    for trial in range(40):
        n = int(rng.integers(5, 400))
        scores = np.round(rng.normal(size=n), 1 if trial % 2 else 6).astype(np.float32)
        labels = rng.integers(0, 2, size=n)
        if labels.min() == labels.max():
            labels[0] = 1 - labels[0]
        two_u, npos, nneg = O.auc_exact(scores, labels)
        brute = sum(2 * (sp > sn) + (sp == sn) for sp in scores[labels == 1]
                    for sn in scores[labels == 0])
        assert two_u == int(brute)
        ref = roc_auc_score(labels, scores)
        assert abs(O.auc_from_counts(two_u, npos, nneg) - ref) <= 4 * np.finfo(np.float64).eps


def test_rank_topk_ties_break_to_lower_index():
    dist = np.array([[3.0, 1.0, 1.0, 0.5, 1.0]])
    vals, idx = O.rank_topk(dist, 3)
    assert idx.tolist() == [[3, 1, 2]] and vals.tolist() == [[0.5, 1.0, 1.0]]


def test_all_pairs_is_the_pair_scorer_on_the_cross_product():
    rng = RNG(17)
    Q, K, d, N = 5, 3, 6, 40
    Pq, E = rng.normal(size=(Q, K, d)), rng.normal(size=(N, d))
    D = O.all_pairs_dist(Pq, E, block=2)
    for q in range(Q):
        np.testing.assert_allclose(D[q], O.pcd_dist(E, np.repeat(Pq[q:q + 1], N, 0)), rtol=1e-12)
    S = T.all_pairs_scores_reference(_t(Pq), _t(E), _t(np.float64(0.7)), pair_batch=7)
    np.testing.assert_allclose(S.numpy(), 0.7 - D, rtol=1e-11, atol=1e-12)
    Sg = T.all_pairs_scores_gram(_t(Pq), _t(E), _t(np.float64(0.7)))
    np.testing.assert_allclose(Sg.numpy(), 0.7 - D, rtol=1e-9, atol=1e-10)


def test_normalisers():
    x = np.array([[0.0, 58.388599, 116.777198]])
    np.testing.assert_allclose(O.normalize(x, 58.388599, 0.0), [[0, 1, 2]])
    y = O.normalize_v2(np.ones((2, 4)), input_shape=(4,), norm=31.9098)
    np.testing.assert_allclose(y, np.full((2, 4), 1 / 31.9098))
    z = O.normalize_v2(np.array([[-3.0, 0.2, 5.0]]), scale=2.0, mean=0.5,
                       clip_value_min=-1.0, clip_value_max=1.0)
    np.testing.assert_allclose(z, [[-1.0, -0.1, 1.0]])
    np.testing.assert_allclose(O.lrelu(np.array([-1.0, 2.0])), [-0.2, 2.0])


def test_adam_tf_first_step_moves_by_lr():
    p, g = np.array([1.0, -2.0]), np.array([0.3, -0.7])
    p1, m, v = O.adam_tf(p, g, np.zeros(2), np.zeros(2), 1, 1e-3)
    np.testing.assert_allclose(p - p1, 1e-3 * np.sign(g), rtol=1e-5)


@pytest.mark.parametrize("name", ["pair_pcd", "pair_modes", "loss", "project", "rank", "rank_monomer", "auc"])
def test_golden_fixtures_reproduce(name):
    """The committed fixtures (tests/golden/make_golden.py) must be reproduced bit-for-bit
    (fp64) by the oracle -- freezes the oracle against accidental edits."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fresh = mg.CASES[name]()
    stored = np.load(os.path.join(GOLD, name + ".npz"))
    assert set(fresh) == set(stored.files)
    for k in fresh:
        np.testing.assert_allclose(np.asarray(fresh[k]), stored[k], rtol=1e-13, atol=0,
                                   err_msg=f"{name}:{k}")


def test_conv2d_weight_norm_oracle_matches_an_independent_convolution():
    """The conv-trunk oracle (cfl/layers.py:100-187: l2_normalize with eps, SAME padding (1, 2) for 28 -> 14) against
    torch's float64 conv2d with explicit asymmetric padding."""
    import torch
    import torch.nn.functional as TF
    from oracle import cfl_oracle as O
    rng = np.random.default_rng(3)
    x = rng.uniform(size=(3, 28, 28, 1))
    V1, g1, b1 = rng.normal(size=(5, 5, 1, 8)), rng.uniform(0.5, 1.5, 8), rng.normal(size=8) * 0.1
    V2, g2, b2 = rng.normal(size=(5, 5, 8, 16)), rng.uniform(0.5, 1.5, 16), rng.normal(size=16) * 0.1
    got = O.conv_pcd_trunk(x.reshape(3, 784), [(V1, g1, b1), (V2, g2, b2)])
    t = torch.as_tensor(x).permute(0, 3, 1, 2)
    for V, g, b in ((V1, g1, b1), (V2, g2, b2)):
        W = torch.as_tensor(V / np.sqrt(np.maximum((V * V).sum((0, 1, 2), keepdims=True), 1e-12)) * g.reshape(1, 1, 1, -1))
        t = TF.conv2d(TF.pad(t, (1, 2, 1, 2)), W.permute(3, 2, 0, 1), bias=torch.as_tensor(b), stride=2)
        t = torch.where(t > 0, t, 0.2 * t)
    want = t.permute(0, 2, 3, 1).reshape(3, -1).numpy()
    assert got.shape == (3, 7 * 7 * 16)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
