"""dV GEMM of the projection (cfl_project_bwd) at the C2 train-step shapes; CFL_EXPERIMENTS=1 CFL_PB_DBG=1|2 isolate the
raw-tile loader / the transposing producers of the TMA-staged kernel (results are garbage then)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat


def ev(fn, n=10):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


g = torch.Generator(device="cuda").manual_seed(1)
for (B, F, N, act) in [(65536, 4096, 100, None), (65536, 4096, 80, None), (65536, 4096, 80, "tanh"), (65536, 1024, 64, None),
                       (8192, 4096, 100, None), (262144, 1024, 192, None)]:
    x = torch.randn(B, F, generator=g, device="cuda")
    V = torch.randn(F, N, generator=g, device="cuda") * 0.02
    b = torch.zeros(N, device="cuda")
    y, _, z = nat.project_fwd(x, V, None, b, False, 1.0, act, want_z=True)
    dy = torch.randn_like(y)
    ms = ev(lambda: nat.project_bwd(x, V, None, b, False, 1.0, act, y, z, dy))
    print(json.dumps(dict(kernel="project_bwd", B=B, F=F, N=N, act=act, ms=round(ms, 4), tflops=round(2.0 * B * F * N / ms / 1e9, 1),
                          x_gbs=round(4.0 * B * F / ms / 1e6, 1), dbg=os.environ.get("CFL_PB_DBG"))), flush=True)
    del x, V, y, z, dy
