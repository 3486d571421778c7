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
if os.environ.get("CFL_PERF_FUSED"):          # only the tensor-core route below
    CASES = []
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

# tensor-core route for pcd: counts inside the fused scoring kernel's epilogue (cfl_rank_counts_packed) beside the exact
# scoring pass of the same kernel family (dense-free: top-1 ranking over the same image) as the "scoring time" yardstick
if not os.environ.get("CFL_PERF_ONLY"):
    for (N, K, d, Q, J) in [(1_000_000, 3, 64, 1024, 8), (1_000_000, 4, 20, 1024, 8), (1_000_000, 1, 64, 1024, 8),
                            (1_000_000, 8, 20, 1024, 8), (1_000_000, 3, 64, 1024, 1)]:
        E = torch.randn(N, d, generator=g, device="cuda")
        Pq = E[torch.randint(0, N, (Q,), generator=g, device="cuda")][:, None, :] + 0.5 * torch.randn(Q, K, d, generator=g, device="cuda")
        mu = nat.col_mean(E)
        img = nat.catalog_pack(E, K, mu)
        pos = torch.randint(0, N, (Q, J), generator=g, device="cuda")
        if os.environ.get("CFL_PERF_PLANTED"):      # labelled positives near the top of the ranking (a trained model)
            D1 = nat.score_topk(Pq, E, 64, mu=mu, image=img)[1]
            pos = D1[:, torch.randperm(64, generator=g, device="cuda")[:J]].contiguous()
        t = nat.pair_dist_rows("pcd", Pq, E, pos)
        got, st = nat.rank_counts_packed(Pq, E, img, mu, t, want_stats=True)
        ms = timeit(lambda: nat.rank_counts_packed(Pq, E, img, mu, t))
        os.environ["CFL_SCORE_NO_LB"] = "1"; os.environ["CFL_SCORE_SAMPLE_STRIDE"] = "1"      # one exact 3xTF32 pass
        ms_s = timeit(lambda: nat.score_topk(Pq, E, 100, mu=mu, image=img))
        del os.environ["CFL_SCORE_NO_LB"], os.environ["CFL_SCORE_SAMPLE_STRIDE"]
        ms_c = timeit(lambda: nat.score_topk(Pq, E, 100, mu=mu, image=img))
        same = None
        if os.environ.get("CFL_PERF_CHECK"):
            same = bool(torch.equal(got, nat.rank_counts("pcd", Pq, E, t)))
        print(json.dumps(dict(kernel="fused_rank_counts", N=N, K=K, d=d, Q=Q, J=J, ms=round(ms, 3), gscores_s=round(Q * N / ms / 1e6, 1),
                              exact_scoring_pass_ms=round(ms_s, 3), cascade_top100_ms=round(ms_c, 3), equals_direct=same,
                              positives="top-64 of the ranking" if os.environ.get("CFL_PERF_PLANTED") else "random rows", **st)), flush=True)
        del E, Pq, img
