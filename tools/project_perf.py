"""Projection kernel: throughput and 3xTF32 error statistics at the benchmark shapes (run under gpurun)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat

def run(B, F, N, reps=5):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, F, generator=g, device="cuda").clamp_(min=0).mul_(10.0)
    V = (torch.rand(F, N, generator=g, device="cuda") * 2 - 1) * (6.0 / (F + N)) ** 0.5
    gg, b = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    for _ in range(2):
        y, _, _ = nat.project_fwd(x, V, gg, b, True, 1 / 31.9098, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y, _, _ = nat.project_fwd(x, V, gg, b, True, 1 / 31.9098, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # error vs fp64 on a row sample
    idx = torch.arange(0, B, max(1, B // 512), device="cuda")[:512]
    xs, Vd = x[idx].double() / 31.9098, V.double()
    want = (xs @ Vd) / Vd.pow(2).sum(0).sqrt()
    bound = (xs.abs() @ Vd.abs()) / Vd.pow(2).sum(0).sqrt()
    err = (y[idx].double() - want).abs()
    return dict(B=B, F=F, N=N, ms=round(ms, 4), items_per_s=B / ms * 1e3, hbm_gbs=B * (4 * F + 4 * N) / ms / 1e6,
                tflops=2.0 * B * F * N / ms / 1e9, max_rel_to_bound=float((err / bound).max()),
                max_rel_to_out=float((err / want.abs().clamp_min(1e-3)).max()))

for shape in [(262144, 1024, 64), (262144, 1024, 192), (131072, 4096, 80), (131072, 2048, 100), (65536, 1024, 256), (100, 4096, 80)]:
    print(json.dumps(run(*shape)), flush=True)
