"""Monomer all-pairs kernel (cfl_score_topk_monomer) throughput at the BASELINE config-2 shapes (run under gpurun).
Bounds: FP32 CUDA-core issue (2 packed FP32x2 instructions per (query pair, catalog float): add + fma) and
catalog streaming N*K*d*4 B per query-tile pass."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = PK.get("hbm_gbs", 6650.0)
SM_MHZ = PK.get("sm_max_mhz", 1965.0)
# packed FP32x2: 148 SMs x 128 lanes x 2 results per instruction per clock
FP32X2_PEAK = 148 * 128 * 2 * SM_MHZ * 1e6


def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks, ke = [], []
    a.record()
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nat.set_kernel_timer(s, e); ks.append(s); ke.append(e)
        fn()
    b.record(); torch.cuda.synchronize()
    nat.set_kernel_timer(None, None)
    return a.elapsed_time(b) / reps, sum(x.elapsed_time(y) for x, y in zip(ks, ke)) / reps


g = torch.Generator(device="cuda").manual_seed(633)
CASES = [(1_000_000, 4, 20, 1024), (1_000_000, 4, 10, 1024), (1_000_000, 4, 20, 64), (1_000_000, 3, 64, 1024),
                     (250_000, 8, 128, 1024)]
if os.environ.get("CFL_PERF_ONLY"):
    CASES = [CASES[int(os.environ["CFL_PERF_ONLY"])]]
for (N, K, d, Q) in CASES:
    P = torch.randn(N, K, d, generator=g, device="cuda")
    a = P[torch.randint(0, N, (Q,), generator=g, device="cuda"), 0] + 0.5 * torch.randn(Q, d, generator=g, device="cuda")
    w = torch.softmax(2 * torch.randn(Q, K, generator=g, device="cuda"), -1)
    ms, kms = timeit(lambda: nat.score_topk_monomer(a, w, P, 100))
    scores = Q * N
    # per score: K*d differences + K*d fmas (2 results per packed instruction) -> 2*K*d flop-equivalents
    print(json.dumps(dict(kernel="score_monomer", N=N, K=K, d=d, Q=Q, ms=round(ms, 3), kernel_ms=round(kms, 3),
                          gscores_s=round(scores / ms / 1e6, 1), kernel_gscores_s=round(scores / kms / 1e6, 1),
                          fp32x2_frac=round(2.0 * K * d * scores / (kms / 1e3) / FP32X2_PEAK, 3),
                          catalog_bytes=N * K * d * 4,
                          hbm_frac_single_read=round(N * K * d * 4 / (kms / 1e3) / 1e9 / HBM, 4))), flush=True)
    mu = P.mean(dim=(0, 1))
    img = nat.monomer_pack(P, mu)
    if img is not None:
        ev, ei = nat.score_topk_monomer(a, w, P, 100)
        tv, ti, st = nat.score_topk_monomer_packed(a, w, P, img, 100, mu=mu, want_stats=True)
        same = bool(torch.equal(ev, tv) and torch.equal(ei, ti))
        ms2, kms2 = timeit(lambda: nat.score_topk_monomer_packed(a, w, P, img, 100, mu=mu))
        dp = (K * (d + 1) + 15) // 16 * 16
        print(json.dumps(dict(kernel="score_monomer_packed (tcgen05 Gram filter + exact rescoring)", N=N, K=K, d=d, Q=Q,
                              ms=round(ms2, 3), filter_kernel_ms=round(kms2, 3), gscores_s=round(scores / ms2 / 1e6, 1),
                              speedup=round(ms / ms2, 2), identical_to_exact_kernel=same,
                              survivors_per_query=round(st[0].item() / Q, 1), redo_queries=int(st[3].item()),
                              filter_tflops=round(2.0 * dp * scores / (kms2 / 1e3) / 1e12, 1),
                              image_bytes=int(img.numel()))), flush=True)
    del P, a, w
