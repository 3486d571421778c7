"""Minimal driver for an ncu capture of pass C (score_lb_kernel) on the bench.py workload: builds the index and issues
three scoring calls.  `ncu --set full --clock-control none --import-source on -k regex:score_lb -s 3 -c 1 python tools/ncu_lb_once.py`
(launch 0/2/4 are the probes, 1/3/5 the full passes: -s 3 skips to the second call's pass C)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
import bench
from cfl import _native as nat
from cfl.ranking import CatalogIndex

dev = torch.device("cuda", 0)
w = bench.synth_weights(dev)
E = torch.empty(bench.N_PER_GPU, bench.D, device=dev)
for lo in range(0, bench.N_PER_GPU, 1 << 18):
    hi = min(bench.N_PER_GPU, lo + (1 << 18))
    xb = bench.synth_features(hi - lo, dev, bench.SEED + 1 + lo // (1 << 18))
    E[lo:hi] = nat.project_fwd(xb, w.V0, w.g0, w.b0, True, w.in_scale, None)[0]
    del xb
index = CatalogIndex(w, E)
xq = bench.synth_features(bench.Q, dev, bench.SEED + 7)
for _ in range(3):
    index.rank(xq, bench.TOPK)
torch.cuda.synchronize()
