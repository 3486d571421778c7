"""Diagnostic for the tcgen05 projection kernel: one-hot feature columns isolate each K-step."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat

B, F, N = 256, 256, 32
rng = np.random.default_rng(0)
V = rng.normal(size=(F, N)).astype(np.float32)
bad = []
for f in list(range(0, 40)) + [63, 64, 100, 255]:
    x = np.zeros((B, F), np.float32); x[:, f] = np.arange(1, B + 1)
    y, _, _ = nat.project_fwd(torch.as_tensor(x).cuda(), torch.as_tensor(V).cuda(), None, None, False, 1.0, None)
    want = x @ V
    err = np.abs(y.cpu().numpy() - want).max() / np.abs(want).max()
    if err > 1e-4:
        got = y.cpu().numpy()
        # which V row does the output look like?
        ratios = got[0] / np.maximum(np.abs(V), 1e-9).T[:, :] if False else None
        cand = [g for g in range(F) if np.allclose(got[0], V[g] * 1.0, rtol=1e-3, atol=1e-4)]
        bad.append((f, float(err), cand[:4], got[0, :3].tolist(), want[0, :3].tolist()))
print("bad one-hot columns:", len(bad))
for b in bad[:20]:
    print(b)
x = rng.normal(size=(B, F)).astype(np.float32)
y, _, _ = nat.project_fwd(torch.as_tensor(x).cuda(), torch.as_tensor(V).cuda(), None, None, False, 1.0, None)
want = x.astype(np.float64) @ V.astype(np.float64)
e = np.abs(y.cpu().numpy() - want)
print("random: max err", e.max(), "rows with err", np.where(e.max(1) > 1e-3)[0][:20], "cols", np.where(e.max(0) > 1e-3)[0][:20])
