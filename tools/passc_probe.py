"""Pass C (score_lb_kernel) on the bench.py workload under the CFL_SCORE_DBG_MODE experiments, in one process:
kernel time of the dominant launch per mode + the survivor statistics of the normal mode.
    0 normal | 1 epilogue does nothing | 3 MMA only (no TMA) | 19 = 3 + MMA issue free-running |
    4 epilogue = tcgen05.ld + wait | 8 = ld + bound, no appends"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
import bench
from cfl import _native as nat
from cfl.ranking import CatalogIndex

os.environ["CFL_EXPERIMENTS"] = "1"     # modes other than 0 need a build with CFL_NVCC_EXTRA=-DCFL_LB_EXPERIMENTS
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
modes = [int(m) for m in sys.argv[1:]] or [0, 1, 3, 19, 4, 8]
for m in modes:
    os.environ["CFL_SCORE_DBG_MODE"] = str(m)
    for _ in range(3):
        index.rank(xq, bench.TOPK)
    ks, ke = [], []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nat.set_kernel_timer(a, b); ks.append(a); ke.append(b)
        index.rank(xq, bench.TOPK)
    e1.record()
    torch.cuda.synchronize()
    nat.set_kernel_timer(None, None)
    kms = sum(a.elapsed_time(b) for a, b in zip(ks, ke)) / len(ks)
    line = f"mode {m:2d}: pass C {kms:.4f} ms, step {e0.elapsed_time(e1) / 20:.4f} ms"
    if m == 0:
        line += f", stats {index.rank_local_stats(xq, bench.TOPK)[2]}"
    print(line, flush=True)
