"""Per-tile timing trace of pass C (CTA (0,0)): who waits for whom.  Needs a trace build:
    CFL_NVCC_EXTRA=-DCFL_LB_TRACE python compatibility-family-learning_b200/build.py --force
    python tools/lb_trace.py [dbg_mode]
(rebuild without the flag afterwards).  Prints the median per-tile durations in clocks."""
import ctypes as C, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
import bench
from cfl import _native as nat
from cfl.ranking import CatalogIndex

os.environ["CFL_EXPERIMENTS"] = "1"
if len(sys.argv) > 1:
    os.environ["CFL_SCORE_DBG_MODE"] = sys.argv[1]
os.environ["CFL_SCORE_NO_PROBE"] = "1"           # the trace must hold the full pass, not the probe
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
n = 8 * 4096
buf = (C.c_ulonglong * n)()
assert nat.lib().cfl_lb_trace_read(buf, n) == 0
tr = np.frombuffer(buf, dtype=np.uint64).reshape(4096, 8).astype(np.int64)
T = 860
tr = tr[:T]
names = ["mma: wait tempty", "mma: issue", "epi: wait tfull", "epi: tcgen05.ld", "epi: compute+append"]
d = [tr[:, 1] - tr[:, 0], tr[:, 2] - tr[:, 1], tr[:, 4] - tr[:, 3], tr[:, 5] - tr[:, 4], tr[:, 6] - tr[:, 5]]
per_tile = np.diff(tr[:, 2])
print("mode", os.environ.get("CFL_SCORE_DBG_MODE", "0"), "clk per tile (MMA issue to issue): median", int(np.median(per_tile)), "mean", int(per_tile.mean()))
for nm, x in zip(names, d):
    x = x[10:]
    print(f"  {nm:22s} median {int(np.median(x)):6d}  mean {int(x.mean()):6d}  p90 {int(np.percentile(x, 90)):6d}")
print("  epilogue tile t starts (tfull seen) after MMA issue of tile t by: median", int(np.median((tr[:, 4] - tr[:, 2])[10:])))
print("  first tiles (relative clk):")
base = tr[0, 0]
for t in range(6):
    print("   ", t, [int(x - base) for x in tr[t, :7]])
