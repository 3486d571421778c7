"""2-GPU NCCL tests (skipped unless two GPUs are visible; run with `gpurun --gpus 2`): catalog
sharding with all-gather + merge kernel, and data-parallel training with gradient all-reduce, both
against the single-GPU result."""
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


def _data(seed=0):
    rng = np.random.default_rng(seed)
    N, F, K, d, Q = 60_000, 64, 3, 32, 50
    X = np.maximum(rng.normal(size=(N, F)), 0).astype(np.float32)
    V0 = ((rng.uniform(size=(F, d)) * 2 - 1) * (6 / (F + d)) ** 0.5).astype(np.float32)
    Vp = ((rng.uniform(size=(F, K * d)) * 2 - 1) * (6 / (F + K * d)) ** 0.5).astype(np.float32)
    return X, V0, Vp, K, d, Q


def _rank_worker(rank, world, port, out):
    _setup(rank, world, port)
    from cfl import ranking
    X, V0, Vp, K, d, Q = _data()
    c = lambda a: torch.as_tensor(a).cuda()
    w = ranking.EncoderWeights(V0=c(V0), Vp=c(Vp), g0=torch.ones(d).cuda(), gp=torch.ones(K * d).cuda(),
                               b0=torch.zeros(d).cuda(), bp=torch.zeros(K * d).cuda())
    lo, hi = ranking.shard_bounds(len(X), world, rank)
    idx = ranking.CatalogIndex.from_features(w, c(X[lo:hi]), idx_base=lo, n_total=len(X))
    tv, ti = idx.rank(c(X[:Q]), 100)
    if rank == 0:
        torch.save(dict(tv=tv.cpu(), ti=ti.cpu(), mu=idx.mu.cpu()), out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_ranking_two_gpus_equals_one(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_rank_worker, args=(2, 29600 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = torch.load(out)
    from cfl import ranking
    X, V0, Vp, K, d, Q = _data()
    c = lambda a: torch.as_tensor(a).cuda()
    w = ranking.EncoderWeights(V0=c(V0), Vp=c(Vp), g0=torch.ones(d).cuda(), gp=torch.ones(K * d).cuda(),
                               b0=torch.zeros(d).cuda(), bp=torch.zeros(K * d).cuda())
    idx = ranking.CatalogIndex.from_features(w, c(X), mu=got["mu"].cuda())      # same centring vector
    tv, ti = idx.rank(c(X[:Q]), 100)
    assert torch.equal(ti.cpu(), got["ti"]) and torch.equal(tv.cpu(), got["tv"])


def _pipelined_worker(rank, world, port, out):
    """rank_async / rank_host (exchange on the side stream over the ring of record buffers, batches in flight beyond
    the ring's depth) must return what the blocking rank() returns for the same batches."""
    _setup(rank, world, port)
    from cfl import ranking
    X, V0, Vp, K, d, Q = _data()
    c = lambda a: torch.as_tensor(a).cuda()
    w = ranking.EncoderWeights(V0=c(V0), Vp=c(Vp), g0=torch.ones(d).cuda(), gp=torch.ones(K * d).cuda(),
                               b0=torch.zeros(d).cuda(), bp=torch.zeros(K * d).cuda())
    lo, hi = ranking.shard_bounds(len(X), world, rank)
    idx = ranking.CatalogIndex.from_features(w, c(X[lo:hi]), idx_base=lo, n_total=len(X))
    k, nb = 100, 6
    xs_host = [torch.as_tensor(X[i * Q:(i + 1) * Q]).pin_memory() for i in range(nb)]
    xs = [x.cuda() for x in xs_host]
    want = [idx.rank(x, k) for x in xs]
    torch.cuda.synchronize()
    got = [idx.rank_async(x, k) for x in xs]                      # six exchanges queued over a ring of four
    for (wv, wi), (gv, gi, ev) in zip(want, got):
        ev.synchronize()
        assert torch.equal(gv, wv) and torch.equal(gi, wi)
    outs = [(torch.empty(Q, k).pin_memory(), torch.empty(Q, k, dtype=torch.int64).pin_memory()) for _ in range(nb)]
    evs = [idx.rank_host(x, k, ov, oi) for x, (ov, oi) in zip(xs_host, outs)]
    for (wv, wi), (ov, oi), ev in zip(want, outs, evs):
        ev.synchronize()
        assert torch.equal(ov, wv.cpu()) and torch.equal(oi, wi.cpu())
    if rank == 0:
        torch.save(dict(ok=True, ti=want[-1][1].cpu()), out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_pipelined_ranking_entry_points_two_gpus(tmp_path):
    out = str(tmp_path / "p.pt")
    mp.spawn(_pipelined_worker, args=(2, 29700 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["ok"] and int(got["ti"].max()) >= 30_000            # global indices reach into the second shard


def _train_worker(rank, world, port, out):
    _setup(rank, world, port)
    from cfl import ops, variables as vs
    from cfl.models.cfl import CFL
    vs.reset_default_graph(); vs.set_seed(633)
    F = 48
    model = CFL(input_shape=(F,), batch_size=64, latent_size=16, num_components=3, model_type="linear",
                dist_type="pcd", use_threshold=True, pos_weight=0.25, reg_const=1e-3, lr=1e-2,
                data_normalizer=ops.normalizer_v2((F,), norm=2.0), data_norm=2.0)
    with torch.no_grad():
        model.raw_threshold.fill_(0.5)
    rng = np.random.default_rng(1)
    for _ in range(3):
        full = [np.maximum(rng.normal(size=(64, F)), 0).astype(np.float32) for _ in range(4)]
        part = [torch.as_tensor(b[rank * 32:(rank + 1) * 32]).cuda() for b in full]
        model.train_step(*part)
    if rank == 0:
        torch.save({k: v.detach().cpu() for k, v in vs.all_variables().items()}, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_training_matches_single_gpu(tmp_path):
    out = str(tmp_path / "w.pt")
    mp.spawn(_train_worker, args=(2, 29700 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = torch.load(out)
    from cfl import ops, variables as vs
    from cfl.models.cfl import CFL
    vs.reset_default_graph(); vs.set_seed(633)
    F = 48
    model = CFL(input_shape=(F,), batch_size=64, latent_size=16, num_components=3, model_type="linear",
                dist_type="pcd", use_threshold=True, pos_weight=0.25, reg_const=1e-3, lr=1e-2,
                data_normalizer=ops.normalizer_v2((F,), norm=2.0), data_norm=2.0)
    with torch.no_grad():
        model.raw_threshold.fill_(0.5)
    rng = np.random.default_rng(1)
    for _ in range(3):
        full = [torch.as_tensor(np.maximum(rng.normal(size=(64, F)), 0).astype(np.float32)).cuda() for _ in range(4)]
        model.train_step(*full)
    for k, v in vs.all_variables().items():
        np.testing.assert_allclose(got[k].numpy(), v.detach().cpu().numpy(), atol=3e-5, rtol=0, err_msg=k)
