"""Minimal driver for an ncu capture of the projection kernels (fwd: project_umma_kernel, bwd: project_bwd_umma_*):
`ncu --set full --clock-control none --import-source on -k regex:project -c 6 python tools/ncu_pb_once.py`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat

g = torch.Generator(device="cuda").manual_seed(1)
B, F, N = 65536, 4096, int(os.environ.get("PB_N", 100))
x = torch.randn(B, F, generator=g, device="cuda")
V = torch.randn(F, N, generator=g, device="cuda") * 0.02
b = torch.zeros(N, device="cuda")
for _ in range(2):
    y, _, z = nat.project_fwd(x, V, None, b, False, 1.0, None, want_z=True)
    dy = torch.randn_like(y)
    nat.project_bwd(x, V, None, b, False, 1.0, None, y, z, dy)
torch.cuda.synchronize()
