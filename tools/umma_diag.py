"""Diagnostic for the tcgen05 building blocks (run under gpurun): prints the error of the
single-tile 3xTF32 GEMM self-test for several shapes and descriptor variants."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(%r, "compatibility-family-learning_b200"))
from cfl import _native as nat
N, Kd = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(N * 1000 + Kd)
A = rng.normal(size=(128, Kd)).astype(np.float32)
B = rng.normal(size=(N, Kd)).astype(np.float32)
D = nat.selftest_umma(torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda())
torch.cuda.synchronize()
D = D.cpu().numpy().astype(np.float64)
want = A.astype(np.float64) @ B.astype(np.float64).T
err = np.abs(D - want)
scale = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64)).T
print("N=%%d Kd=%%d max_abs_err=%%.3e max_rel_to_scale=%%.3e  (1xTF32 would be ~1e-3)  D[0,:4]=%%s want=%%s" %% (
    N, Kd, err.max(), (err / scale).max(), D[0, :4], want[0, :4]))
''' % ROOT

for variant in sys.argv[1:] or ["0"]:
    for N, Kd in [(16, 8), (64, 8), (64, 16), (192, 64), (256, 64), (112, 24), (64, 20)]:
        env = dict(os.environ, CFL_UMMA_VARIANT=variant)
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, str(N), str(Kd)], env=env, capture_output=True,
                               text=True, timeout=120)
            out = (r.stdout.strip() or r.stderr.strip()[-400:])
        except subprocess.TimeoutExpired:
            out = "TIMEOUT"
        print(f"[variant {variant}] {out}", flush=True)
