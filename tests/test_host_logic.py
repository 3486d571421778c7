"""CPU tests of the host-side mirror: flag parsers, naming, normaliser folding, variable scopes,
and the N>1 sharded ranking path on a world-size-2 gloo group (kernels replaced by oracle-backed
test doubles -- the product itself has no CPU path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cfl_oracle as O


def test_parsers_keep_reference_defaults():
    from cfl import utils
    a = utils.monomer_parser().parse_args([])
    assert (a.num_components, a.latent_size, a.lr, a.beta1, a.beta2, a.seed) == (2, 20, 0.001, 0.9, 0.999, 633)
    assert tuple(a.input_shape) == (4096,) and a.normalize_value == 1.0
    b = utils.dist_parser().parse_args(["--dist-type", "pcd", "--use-threshold", "--pos-weight", "0.0625",
                                       "--data-norm", "31.9098", "--latent-size", "64", "--num-components", "3"])
    utils.dist_check_args(b)
    assert b.data_norm == (31.9098,) and b.pos_weight == 0.0625 and b.model_type == "conv"
    with pytest.raises(AssertionError):
        utils.dist_check_args(utils.dist_parser().parse_args(["--dist-type", "pcd"]))      # needs --use-threshold
    with pytest.raises(AssertionError):
        utils.dist_check_args(utils.dist_parser().parse_args(["--dist-type", "pcd", "--use-threshold", "--caffe-margin", "1"]))


def test_normalisers_fold_into_in_scale_only_when_pure():
    from cfl import ops
    assert abs(ops.normalizer(58.388599, 0.0).in_scale - 1 / 58.388599) < 1e-12
    assert ops.normalizer(2.0, 0.5).in_scale is None
    assert abs(ops.normalizer_v2((4,), norm=31.9098).in_scale - 1 / 31.9098) < 1e-12
    assert ops.normalizer_v2((4,), norm=2.0, clip_value_min=0.0).in_scale is None
    x = torch.tensor([[-3.0, 0.2, 5.0]])
    np.testing.assert_allclose(ops.normalize_v2(x, None, scale=2.0, mean=0.5, clip_value_min=-1.0, clip_value_max=1.0),
                               O.normalize_v2(x.numpy(), None, scale=2.0, mean=0.5, clip_value_min=-1.0, clip_value_max=1.0))
    np.testing.assert_allclose(ops.lrelu(torch.tensor([-1.0, 2.0])), [-0.2, 2.0])
    dn = ops.dist_normalizer((4,), None, None, None, 31.9098, None, "linear")[0]
    np.testing.assert_allclose(dn(torch.ones(2, 4)), np.full((2, 4), 1 / 31.9098), rtol=1e-6)


def test_variable_scope_reuse_semantics():
    from cfl import variables as vs
    vs.reset_default_graph()
    vs.set_default_device("cpu")
    try:
        with vs.variable_scope("enc") as sc:
            with vs.variable_scope("outputs"):
                v = vs.get_variable("V", [3, 2], vs.xavier_initializer())
        assert sc.name == "enc" and "enc/outputs/V" in vs.all_variables()
        with pytest.raises(ValueError):
            with vs.variable_scope("enc"), vs.variable_scope("outputs"):
                vs.get_variable("V", [3, 2], vs.xavier_initializer())
        with vs.variable_scope("enc", reuse=True), vs.variable_scope("outputs"):
            assert vs.get_variable("V", [3, 2], vs.xavier_initializer()) is v
        with pytest.raises(ValueError):
            with vs.variable_scope("enc", reuse=True):
                vs.get_variable("missing", [1], vs.zeros_initializer())
        with vs.variable_scope(sc), vs.variable_scope("t"):           # re-enter a captured scope
            vs.get_variable("threshold", [], vs.constant_initializer(1e-6))
        assert float(vs.all_variables()["enc/t/threshold"].detach()) == pytest.approx(1e-6)
    finally:
        vs.set_default_device(None)
        vs.reset_default_graph()


def test_model_names_match_reference_format():
    from cfl.models.cfl import CFL
    from cfl.models.dist import Dist
    m = CFL.__new__(CFL)
    m.__dict__.update(dist_type="pcd", model_type="linear", directed=False, pos_weight=0.0625, caffe_margin=None,
                      data_type="linear", latent_size=64, num_components=3, act_type=None, use_threshold=True,
                      reg_const=0.0, data_norm=(31.9098,), lambda_m=None, run_tag=None)
    assert m.get_name() == "cfl_pcd_linear_pw_0.0625_linear_ls_64_nc_3_ut_norm_31.9098"     # experiments/dyadic/eval.sh:15
    d = Dist.__new__(Dist)
    d.__dict__.update(latent_size=20, num_components=4, reg_const=0.0, normalize_value=58.388599, run_tag=None)
    assert d.get_name() == "linear_dist_ls_20_nc_4_reg_0.0_norm_58.388599"


def test_shard_bounds_partition_the_catalog():
    from cfl.ranking import shard_bounds
    for n, w in [(10, 3), (1_000_000, 8), (7, 8)]:
        cuts = [shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


# ---------------------------------------------------------------------------------------------
def _doubles():
    """Oracle-backed stand-ins for the kernels the ranking path calls (CPU, test only)."""
    from cfl import _native as nat

    def project_fwd(x, V, g=None, bias=None, weight_norm=True, in_scale=1.0, act=None, **kw):
        y = O.fc_weight_norm(x.numpy().astype(np.float64) * in_scale, V.numpy().astype(np.float64),
                             None if g is None else g.numpy().astype(np.float64),
                             None if bias is None else bias.numpy().astype(np.float64), act)
        return torch.as_tensor(y, dtype=torch.float32), None, None

    def score_topk(Pq, E, k, mu=None, mode="pcd", idx_base=0, want_dense=False, image=None):
        D = O.all_pairs_dist(Pq.numpy().astype(np.float64), E.numpy().astype(np.float64))
        v, i = O.rank_topk(D, min(k, E.shape[0]))
        return torch.as_tensor(v, dtype=torch.float32), torch.as_tensor(i + idx_base)

    def topk_merge(vals, idx):
        R, Q, k = vals.shape
        v = vals.permute(1, 0, 2).reshape(Q, R * k).numpy()
        i = idx.permute(1, 0, 2).reshape(Q, R * k).numpy()
        order = np.lexsort((i, v), axis=1)[:, :k]
        return torch.as_tensor(np.take_along_axis(v, order, 1)), torch.as_tensor(np.take_along_axis(i, order, 1))

    nat.project_fwd, nat.score_topk, nat.topk_merge = project_fwd, score_topk, topk_merge
    nat.col_mean = lambda E: E.double().mean(0).float()
    nat.catalog_pack = lambda E, K, mu=None: None


def _rank_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "compatibility-family-learning_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cfl import ranking
    _doubles()
    ranking.CatalogIndex.__init__.__globals__["nat"].CflNativeError = RuntimeError
    rng = np.random.default_rng(0)
    N, F, K, d, Q, k = 401, 12, 3, 6, 9, 20                       # 401 rows: ragged shards at every world size
    X = rng.normal(size=(N, F)).astype(np.float32)
    w = ranking.EncoderWeights(V0=torch.as_tensor(O.xavier_uniform(rng, F, d)), Vp=torch.as_tensor(O.xavier_uniform(rng, F, K * d)),
                               g0=torch.ones(d), gp=torch.ones(K * d), b0=torch.zeros(d), bp=torch.zeros(K * d))
    lo, hi = ranking.shard_bounds(N, world, rank)
    E = ranking.nat.project_fwd(torch.as_tensor(X[lo:hi]), w.V0, w.g0, w.b0)[0]
    idx = ranking.CatalogIndex.__new__(ranking.CatalogIndex)      # bypass the CUDA-only check of __init__
    idx.w, idx.E, idx.idx_base, idx.n_total, idx.group, idx.theta, idx.image = w, E, lo, N, None, 1.0, None
    idx.mu = idx._global_mean()
    tv, ti = idx.rank(torch.as_tensor(X[:Q]), k)
    if rank == 0:
        np.savez(os.path.join(out_dir, "r.npz"), tv=tv.numpy(), ti=ti.numpy(), mu=idx.mu.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ranking_world2_gloo_matches_single_process(tmp_path, world):
    """Catalog rows sharded over 2 / 4 ranks (ragged shards) + one all_gather of the records + merge == ranking the
    whole catalog: the result does not depend on the number of ranks."""
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_rank_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "r.npz")
    rng = np.random.default_rng(0)
    N, F, K, d, Q, k = 401, 12, 3, 6, 9, 20
    X = rng.normal(size=(N, F)).astype(np.float32)
    V0, Vp = O.xavier_uniform(rng, F, d), O.xavier_uniform(rng, F, K * d)
    E = O.fc_weight_norm(X.astype(np.float64), V0.astype(np.float64), np.ones(d), np.zeros(d)).astype(np.float32)
    P = O.fc_weight_norm(X[:Q].astype(np.float64), Vp.astype(np.float64), np.ones(K * d), np.zeros(K * d)).astype(np.float32)
    D = O.all_pairs_dist(P.reshape(Q, K, d).astype(np.float64), E.astype(np.float64))
    wv, wi = O.rank_topk(D, k)
    assert (got["ti"] == wi).all()
    np.testing.assert_allclose(got["tv"], wv, rtol=1e-6)
    np.testing.assert_allclose(got["mu"], E.astype(np.float64).mean(0), rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------
def _auc_counts_double(pos, neg):
    """Oracle-backed stand-in for cfl_auc (CPU, test only): [twoU, n_pos, n_neg, correct@0]."""
    s = np.concatenate([pos.numpy(), neg.numpy()])
    y = np.concatenate([np.ones(pos.numel(), np.uint8), np.zeros(neg.numel(), np.uint8)])
    two_u, n_pos, n_neg = O.auc_exact(s, y)
    correct = int((pos > 0).sum()) + int((neg <= 0).sum())
    return torch.tensor([two_u, n_pos, n_neg, correct], dtype=torch.int64)


def _pair_scores():
    rng = np.random.default_rng(21)
    pos = np.round(rng.normal(size=1001) + 0.4, 1).astype(np.float32)     # rounded: many exact ties
    neg = np.round(rng.normal(size=4003), 1).astype(np.float32)
    return pos, neg


def _sharded_auc_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "compatibility-family-learning_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cfl import _native as nat
    from cfl import utils
    nat.auc_counts = _auc_counts_double
    pos, neg = _pair_scores()
    # uneven shards; rank 1 holds no positives at all in the second case
    a = utils.sharded_auc_counts(torch.as_tensor(pos[rank::world]), torch.as_tensor(neg[rank::world]))
    b = utils.sharded_auc_counts(torch.as_tensor(pos if rank == 0 else pos[:0]), torch.as_tensor(neg[rank * 3000:(rank + 1) * 3000]))
    if rank == 1:
        np.savez(os.path.join(out_dir, "a.npz"), a=np.array(a, dtype=np.int64), b=np.array(b, dtype=np.int64))
    dist.destroy_process_group()


def test_sharded_pair_auc_world2_gloo_equals_single_process(tmp_path):
    """labelled-pair scores sharded over ranks -> the same AUC / accuracy integers as one process."""
    port = 33500 + os.getpid() % 2000
    mp.spawn(_sharded_auc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "a.npz")
    pos, neg = _pair_scores()
    want = _auc_counts_double(torch.as_tensor(pos), torch.as_tensor(neg)).tolist()
    assert got["a"].tolist() == want and got["b"].tolist() == want


def test_unnormalisers_invert_the_normalisers():
    """cfl/ops.py:146-198,217-225: unnormalize(_v2) undo normalize(_v2) (no clip), also per channel."""
    from cfl import ops
    g = torch.Generator().manual_seed(5)
    x = torch.rand(6, 4, 4, 3, generator=g, dtype=torch.float64) * 255
    kw = dict(scale=1 / 255.0, mean=(0.4, 0.5, 0.6), norm=(0.2, 0.3, 0.25))
    y = ops.normalize_v2(x, (4, 4, 3), **kw)
    assert y.shape == (6, 48)
    back = ops.unnormalizer_v2((4, 4, 3), **kw)(y)
    assert back.shape == x.shape
    np.testing.assert_allclose(back.numpy(), x.numpy(), rtol=1e-12, atol=1e-9)
    y1 = ops.normalize_v2(x[..., 0], (4, 4), scale=2.0, mean=0.5, norm=31.9098)
    np.testing.assert_allclose(ops.unnormalize_v2(y1, (4, 4), scale=2.0, mean=0.5, norm=31.9098).numpy(), x[..., 0].numpy(),
                               rtol=1e-12)
    z = ops.normalize(x, 58.388599, 0.25)
    np.testing.assert_allclose(ops.unnormalizer(58.388599, 0.25)(z).numpy(), x.numpy(), rtol=1e-12, atol=1e-9)
    dn, dun, an, aun, ln = ops.dist_normalizer((4, 4, 3), (2, 2, 3), 1 / 255.0, None, (0.5,), 7.0, "linear")
    np.testing.assert_allclose(dun(dn(x)).numpy(), x.numpy(), rtol=1e-12, atol=1e-9)
    assert an is not dn and aun is not dun and ln is not None
    dn2, dun2, an2, aun2, ln2 = ops.dist_normalizer((4,), None, None, None, 31.9098, None, "linear")
    assert an2 is dn2 and aun2 is dun2 and ln2 is None


# ---------------------------------------------------------------------------------------------
def _mono_doubles(nat):
    """Oracle-backed stand-ins for the kernels the monomer ranking path calls (CPU, test only)."""
    def project_fwd(x, V, g=None, bias=None, weight_norm=True, in_scale=1.0, act=None, want_pre=False, **kw):
        y, pre = O.fc_weight_norm(x.numpy().astype(np.float64) * in_scale, V.numpy().astype(np.float64),
                                  None if g is None else g.numpy().astype(np.float64),
                                  None if bias is None else bias.numpy().astype(np.float64), act, return_pre=True)[:2]
        return torch.as_tensor(y, dtype=torch.float32), torch.as_tensor(pre, dtype=torch.float32), None

    def score_topk_monomer(a, w, P, k, idx_base=0, want_dense=False):
        D = O.all_pairs_monomer_dist(a.numpy().astype(np.float64), w.numpy().astype(np.float64), P.numpy().astype(np.float64))
        v, i = O.rank_topk(D, min(k, P.shape[0]))
        return torch.as_tensor(v, dtype=torch.float32), torch.as_tensor(i + idx_base)

    def topk_merge(vals, idx):
        R, Q, k = vals.shape
        v = vals.permute(1, 0, 2).reshape(Q, R * k).numpy()
        i = idx.permute(1, 0, 2).reshape(Q, R * k).numpy()
        order = np.lexsort((i, v), axis=1)[:, :k]
        return torch.as_tensor(np.take_along_axis(v, order, 1)), torch.as_tensor(np.take_along_axis(i, order, 1))

    nat.project_fwd, nat.score_topk_monomer, nat.topk_merge = project_fwd, score_topk_monomer, topk_merge


def _mono_case():
    rng = np.random.default_rng(31)
    N, F, K, d, Q, k = 300, 10, 3, 5, 8, 15
    X = rng.normal(size=(N, F)).astype(np.float32)
    return X, O.xavier_uniform(rng, F, d), O.xavier_uniform(rng, F, K * d), O.xavier_uniform(rng, d, K), K, d, Q, k


def _mono_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "compatibility-family-learning_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cfl import ranking
    _mono_doubles(ranking.nat)
    X, V0, Vp, Vg, K, d, Q, k = _mono_case()
    t = torch.as_tensor
    w = ranking.EncoderWeights(V0=t(V0), Vp=t(Vp), g0=torch.ones(d), gp=torch.ones(K * d), Vg=t(Vg), gg=torch.ones(K), act="tanh")
    lo, hi = ranking.shard_bounds(len(X), world, rank)
    P = ranking.nat.project_fwd(t(X[lo:hi]), w.Vp, w.gp, None, True, 1.0, "tanh")[0]
    idx = ranking.MonomerCatalogIndex.__new__(ranking.MonomerCatalogIndex)     # bypass the CUDA-only check of __init__
    idx.w, idx.P, idx.idx_base, idx.n_total, idx.group, idx.theta = w, P.view(-1, K, d), lo, len(X), None, 1.0
    idx.image, idx.mu = None, idx._global_mean()              # (no tensor-core image on the CPU; the mean is a collective)
    tv, ti = idx.rank(t(X[:Q]), k)
    if rank == 1:
        np.savez(os.path.join(out_dir, "m.npz"), tv=tv.numpy(), ti=ti.numpy())
    dist.destroy_process_group()


def test_sharded_monomer_ranking_world2_gloo_matches_single_process(tmp_path):
    """MonomerCatalogIndex: target prototypes sharded over 2 ranks + all_gather + merge == the whole catalog; the
    query side (embedding + gate softmax of the pre-activation embedding, base.py:94-105) follows the oracle."""
    port = 35500 + os.getpid() % 2000
    mp.spawn(_mono_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "m.npz")
    X, V0, Vp, Vg, K, d, Q, k = _mono_case()
    f = lambda a: a.astype(np.float64)
    params = {"outputs": (f(V0), np.ones(d), None), "prototype_outputs": (f(Vp), np.ones(K * d), None),
              "monomer_outputs": (f(Vg), np.ones(K), None)}
    S = O.build_prototypes(f(X[:Q]), params, "monomer", K, d, "tanh")
    Tt = O.build_prototypes(f(X), params, "monomer", K, d, "tanh")
    D = O.all_pairs_monomer_dist(S["activations"], S["monomer_activations"], Tt["prototype_activations"])
    wv, wi = O.rank_topk(D, k)
    assert (got["ti"] == wi).all()
    np.testing.assert_allclose(got["tv"], wv, rtol=1e-5)
