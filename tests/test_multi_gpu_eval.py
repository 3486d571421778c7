"""2-GPU NCCL tests of the evaluation side (skipped unless two GPUs are visible; run with `gpurun --gpus 2`):
per-query all-candidate AUC (direct and tensor-core routes), monomer ranking and labelled-pair AUC with the
catalog / the pairs sharded over two ranks -- every integer equal to the single-GPU result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "compatibility-family-learning_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _data(seed=3):
    rng = np.random.default_rng(seed)
    N, F, K, d, Q, J = 50_000, 64, 3, 32, 40, 6
    X = np.maximum(rng.normal(size=(N, F)), 0).astype(np.float32)
    xav = lambda a, b: ((rng.uniform(size=(a, b)) * 2 - 1) * (6 / (a + b)) ** 0.5).astype(np.float32)
    V0, Vp, Vg = xav(F, d), xav(F, K * d), xav(d, K)
    pos = rng.permutation(np.arange(Q, N))[:Q * J].reshape(Q, J).astype(np.int64)
    pos[::4, -1] = -1
    ps = np.round(rng.normal(size=3001) + 0.4, 1).astype(np.float32)
    ns = np.round(rng.normal(size=9005), 1).astype(np.float32)
    return X, V0, Vp, Vg, K, d, Q, pos, ps, ns


def _weights(V0, Vp, Vg, K, d):
    from cfl import ranking
    c = lambda a: torch.as_tensor(a).cuda()
    return ranking.EncoderWeights(V0=c(V0), Vp=c(Vp), g0=torch.ones(d).cuda(), gp=torch.ones(K * d).cuda(),
                                  b0=torch.zeros(d).cuda(), bp=torch.zeros(K * d).cuda(), Vg=c(Vg), gg=torch.ones(K).cuda())


def _evaluate(world, rank, mu=None):
    from cfl import ranking, utils
    X, V0, Vp, Vg, K, d, Q, pos, ps, ns = _data()
    c = lambda a: torch.as_tensor(a).cuda()
    w = _weights(V0, Vp, Vg, K, d)
    lo, hi = ranking.shard_bounds(len(X), world, rank)
    idx = ranking.CatalogIndex.from_features(w, c(X[lo:hi]), idx_base=lo, n_total=len(X), mu=mu)
    a = idx.auc_per_query(c(X[:Q]), torch.as_tensor(pos), method="direct")
    g = idx.auc_per_query(c(X[:Q]), torch.as_tensor(pos), method="gram")
    mono = ranking.MonomerCatalogIndex.from_features(w, c(X[lo:hi]), idx_base=lo, n_total=len(X))
    mv, mi = mono.rank(c(X[:Q]), 50)
    ma = mono.auc_per_query(c(X[:Q]), torch.as_tensor(pos))
    pair = utils.sharded_auc_counts(c(ps[rank::world]), c(ns[rank::world]))
    return dict(mu=idx.mu.cpu(), direct=a.counts.cpu(), direct_t=a.pos_dist.cpu(), gram=g.counts.cpu(), gram_t=g.pos_dist.cpu(),
                mv=mv.cpu(), mi=mi.cpu(), mono=ma.counts.cpu(), mono_two_u=ma.two_u.cpu(), pair=torch.tensor(pair))


def _eval_worker(rank, world, port, out):
    _setup(rank, world, port)
    r = _evaluate(world, rank)
    if rank == 1:                       # any rank holds the full result
        torch.save(r, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_evaluation_two_gpus_equals_one(tmp_path):
    out = str(tmp_path / "e.pt")
    mp.spawn(_eval_worker, args=(2, 29800 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = torch.load(out)
    for p in (ROOT, os.path.join(ROOT, "compatibility-family-learning_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    want = _evaluate(1, 0, mu=got["mu"].cuda())          # same centring vector as the sharded run
    for k in ("direct", "direct_t", "gram", "gram_t", "mi", "mv", "mono", "mono_two_u", "pair"):
        a, b = got[k], want[k]
        if a.is_floating_point():
            a, b = torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)
        assert torch.equal(a, b), k
