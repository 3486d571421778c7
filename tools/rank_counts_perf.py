"""Per-query all-candidate AUC kernels (cfl_pair_dist_rows + cfl_rank_counts) at the C3 / C2 shapes (run under gpurun)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat


def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


g = torch.Generator(device="cuda").manual_seed(633)
CASES = [("pcd", 1_000_000, 3, 64, 1024, 8), ("pcd", 1_000_000, 4, 20, 1024, 8), ("pcd", 1_000_000, 1, 64, 1024, 8),
                              ("monomer", 1_000_000, 4, 20, 1024, 8)]
if os.environ.get("CFL_PERF_ONLY"):
    CASES = [CASES[int(os.environ["CFL_PERF_ONLY"])]]
for (mode, N, K, d, Q, J) in CASES:
    pos = torch.randint(0, N, (Q, J), generator=g, device="cuda")
    if mode == "pcd":
        cat = torch.randn(N, d, generator=g, device="cuda")
        qry = cat[torch.randint(0, N, (Q,), generator=g, device="cuda")][:, None, :] + 0.5 * torch.randn(Q, K, d, generator=g, device="cuda")
        w = None
    else:
        cat = torch.randn(N, K, d, generator=g, device="cuda")
        qry = cat[torch.randint(0, N, (Q,), generator=g, device="cuda"), 0] + 0.5 * torch.randn(Q, d, generator=g, device="cuda")
        w = torch.softmax(torch.randn(Q, K, generator=g, device="cuda"), -1)
    t = nat.pair_dist_rows(mode, qry, cat, pos, w=w)
    ms_t = timeit(lambda: nat.pair_dist_rows(mode, qry, cat, pos, w=w))
    ms = timeit(lambda: nat.rank_counts(mode, qry, cat, t, w=w))
    print(json.dumps(dict(kernel="rank_counts", mode=mode, N=N, K=K, d=d, Q=Q, J=J, pos_dist_ms=round(ms_t, 3), ms=round(ms, 3),
                          gscores_s=round(Q * N / ms / 1e6, 1))), flush=True)
    del cat, qry

# tensor-core route for pcd: dense Gram-form distances from the scoring kernel + HBM-speed counting
if not os.environ.get("CFL_PERF_ONLY"):
    for (N, K, d, Q, J) in [(1_000_000, 3, 64, 1024, 8), (1_000_000, 4, 20, 1024, 8)]:
        E = torch.randn(N, d, generator=g, device="cuda")
        Pq = E[torch.randint(0, N, (Q,), generator=g, device="cuda")][:, None, :] + 0.5 * torch.randn(Q, K, d, generator=g, device="cuda")
        mu = nat.col_mean(E)
        img = nat.catalog_pack(E, K, mu)
        pos = torch.randint(0, N, (Q, J), generator=g, device="cuda")
        dense = nat.score_topk(Pq, E, 1, mu=mu, image=img, want_dense=True)[2]
        t = torch.gather(dense, 1, pos)
        ms_d = timeit(lambda: nat.score_topk(Pq, E, 1, mu=mu, image=img, want_dense=True))
        ms_c = timeit(lambda: nat.dense_rank_counts(dense, t))
        print(json.dumps(dict(kernel="gram_route", N=N, K=K, d=d, Q=Q, J=J, dense_scoring_ms=round(ms_d, 3), count_ms=round(ms_c, 3),
                              count_hbm_gbs=round(Q * N * 4 / ms_c / 1e6, 1), gscores_s=round(Q * N / (ms_d + ms_c) / 1e6, 1))), flush=True)
        del E, Pq, dense, img
