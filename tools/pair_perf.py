"""Paired distance/loss kernels and the AUC kernel vs the HBM roofline (run under gpurun)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

g = torch.Generator(device="cuda").manual_seed(1)
th = torch.tensor([1.0], device="cuda")
for (B, K, d) in [(4_000_000, 3, 64), (8_000_000, 4, 20), (8_000_000, 4, 10), (1_000_000, 8, 128), (100, 4, 20)]:
    v = torch.randn(B, d, generator=g, device="cuda")
    P = torch.randn(B, K, d, generator=g, device="cuda")
    ms_f = timeit(lambda: nat.pair_loss_fwd("pcd", v, P, theta=th, label=1, want_score=True, want_stats=True))
    da = torch.empty(B, d, device="cuda"); dP = torch.empty(B, K * d, device="cuda")
    ms_b = timeit(lambda: nat.pair_loss_bwd("pcd", v, P, theta=th, label=1, c_ce=1.0 / B, want_dtheta=True, da=da, dP=dP))
    bf = B * (4 * d * (K + 1) + 8)
    bb = B * (2 * 4 * d * (K + 1))
    print(json.dumps(dict(kernel="pair", B=B, K=K, d=d, fwd_ms=round(ms_f, 4), fwd_gpairs_s=round(B / ms_f / 1e6, 2),
                          fwd_hbm_frac=round(bf / ms_f / 1e6 / HBM, 3), bwd_ms=round(ms_b, 4),
                          bwd_gpairs_s=round(B / ms_b / 1e6, 2), bwd_hbm_frac=round(bb / ms_b / 1e6 / HBM, 3))), flush=True)
    del v, P, da, dP
for (npos, nneg) in [(125_000, 2_000_000), (1_000_000, 16_000_000)]:
    pos = torch.randn(npos, generator=g, device="cuda") + 0.3
    neg = torch.randn(nneg, generator=g, device="cuda")
    ms = timeit(lambda: nat.auc_counts(pos, neg), reps=5)
    print(json.dumps(dict(kernel="auc", n_pos=npos, n_neg=nneg, ms=round(ms, 4), mpairs_scored_per_s=round((npos + nneg) / ms / 1e3, 1),
                          hbm_frac=round((nneg * 4 * 13 + npos * 4) / ms / 1e6 / HBM, 3))), flush=True)
