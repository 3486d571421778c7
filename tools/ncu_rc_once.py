"""Minimal driver for an ncu capture of the fused rank-count kernel (rank_count_umma_kernel) at the C3 shape:
`ncu --set full --clock-control none --import-source on -k regex:rank_count_umma -s 1 -c 1 python tools/ncu_rc_once.py`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat

g = torch.Generator(device="cuda").manual_seed(633)
N, K, d, Q, J = 1_000_000, int(os.environ.get("RC_K", 3)), int(os.environ.get("RC_D", 64)), 1024, 8
E = torch.randn(N, d, generator=g, device="cuda")
Pq = E[torch.randint(0, N, (Q,), generator=g, device="cuda")][:, None, :] + 0.5 * torch.randn(Q, K, d, generator=g, device="cuda")
mu = nat.col_mean(E)
img = nat.catalog_pack(E, K, mu)
pos = torch.randint(0, N, (Q, J), generator=g, device="cuda")
t = nat.pair_dist_rows("pcd", Pq, E, pos)
for _ in range(3):
    nat.rank_counts_packed(Pq, E, img, mu, t)
torch.cuda.synchronize()
