"""One shape of the K x d sweep with the lower-bound pass on / off and the debug counters."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat
K, d = int(sys.argv[1]), int(sys.argv[2])
N, Q, k = 1_250_000, 1024, 100
g = torch.Generator(device="cuda").manual_seed(633)
E = torch.randn(N, d, generator=g, device="cuda"); mu = nat.col_mean(E)
anchors = torch.randint(0, N, (Q,), generator=g, device="cuda")
Pq = (E[anchors][:, None, :] + 0.5 * torch.randn(Q, K, d, generator=g, device="cuda")).contiguous()
img = nat.catalog_pack(E, K, mu)
for _ in range(2): nat.score_topk(Pq, E, k, mu=mu, image=img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): nat.score_topk(Pq, E, k, mu=mu, image=img)
e1.record(); torch.cuda.synchronize()
print(json.dumps(dict(K=K, d=d, env={k_: v for k_, v in os.environ.items() if k_.startswith("CFL_")}, ms=round(e0.elapsed_time(e1) / 5, 3))), flush=True)
nat.score_topk(Pq, E, k, mu=mu, image=img); torch.cuda.synchronize()
