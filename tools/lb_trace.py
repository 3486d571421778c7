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
T = 1100
tr = tr[20:T]
names = ["epi: wait tfull", "epi: tcgen05.ld + wait", "epi: release + math", "epi: ballot + hit path", "epi: loop back to next wait"]
d = [tr[:, 4] - tr[:, 3], tr[:, 5] - tr[:, 4], tr[:, 6] - tr[:, 5], tr[:, 7] - tr[:, 6], tr[1:, 3] - tr[:-1, 7]]
per = np.diff(tr[:, 4])
print("mode", os.environ.get("CFL_SCORE_DBG_MODE", "0"), "clk per step (epilogue warp 0, tfull seen to tfull seen): median", int(np.median(per)), "mean", int(per.mean()))
for nm, x in zip(names, d):
    print(f"  {nm:28s} median {int(np.median(x)):6d}  mean {int(x.mean()):6d}  p90 {int(np.percentile(x, 90)):6d}")
hitp = tr[:, 7] - tr[:, 6]
print("  steps with a hit path > 60 clk:", float((hitp > 60).mean()), " mean cost of those:", int(hitp[hitp > 60].mean()) if (hitp > 60).any() else 0)
