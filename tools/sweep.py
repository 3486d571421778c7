"""BASELINE config 5 sweep on one GPU: fused scoring + top-100 for K in {1,2,4,8} x d in {20,64,128},
catalog = 1.25M rows (a 10M catalog over 8 GPUs), Q = 1024.  Embeddings generated directly
(e ~ N(0,1)^d, p_k = e_anchor + 0.5 N(0,1)).  Prints one JSON line per shape."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat

PEAK_TF32 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] / 2 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 700.0
N, Q, k = 1_250_000, 1024, 100
g = torch.Generator(device="cuda").manual_seed(633)
for d in (20, 64, 128):
    E = torch.randn(N, d, generator=g, device="cuda")
    mu = nat.col_mean(E)
    for K in (1, 2, 4, 8):
        anchors = torch.randint(0, N, (Q,), generator=g, device="cuda")
        Pq = (E[anchors][:, None, :] + 0.5 * torch.randn(Q, K, d, generator=g, device="cuda")).contiguous()
        img = nat.catalog_pack(E, K, mu)
        for _ in range(2):
            nat.score_topk(Pq, E, k, mu=mu, image=img)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            tv, ti = nat.score_topk(Pq, E, k, mu=mu, image=img)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ok = bool((ti[:, 0] == anchors).float().mean() > 0.5) if K >= 1 else True
        tfl = 2.0 * K * d * Q * N / ms / 1e9
        print(json.dumps(dict(K=K, d=d, N=N, Q=Q, ms=round(ms, 3), gscores_per_s=round(Q * N / ms / 1e6, 1),
                              tflops=round(tfl, 1), frac_tf32=round(tfl / PEAK_TF32, 4), frac_3xtf32=round(3 * tfl / PEAK_TF32, 4),
                              tensor_path=img is not None)), flush=True)
        del img
