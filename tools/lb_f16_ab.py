"""A/B of the fp16 operand planes of the lower-bound pass (CFL_SCORE_LB_F16) on the bench.py workload, in one
process: identical top-100 lists required, then the time of the scoring call with and without (run under gpurun)."""
import json, os, sys
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
Pq = [index.project_queries(bench.synth_features(bench.Q, dev, bench.SEED + 7 + i)) for i in range(4)]


def run(reps=24):
    for i in range(3): index.rank_local(Pq[i % 4], bench.TOPK)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks, ke = [], []
    a.record()
    for i in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nat.set_kernel_timer(s, e); ks.append(s); ke.append(e)
        index.rank_local(Pq[i % 4], bench.TOPK)
    b.record(); torch.cuda.synchronize()
    nat.set_kernel_timer(None, None)
    return a.elapsed_time(b) / reps, sum(x.elapsed_time(y) for x, y in zip(ks, ke)) / reps


out = {}
for mode in ("0", "1", "0", "1"):
    os.environ["CFL_SCORE_LB_F16"] = mode
    res = [index.rank_local(p, bench.TOPK) for p in Pq]
    ms, kms = run()
    if mode in out:
        same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(res, out[mode]))
    else:
        out[mode] = res
        same = None
    print(json.dumps(dict(lb_f16=int(mode), ms=round(ms, 4), pass_c_ms=round(kms, 4), deterministic=same)), flush=True)
same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(out["0"], out["1"]))
print(json.dumps(dict(f16_equals_tf32_results=same)), flush=True)
# the range guard: a catalog with a huge centred value must fall back to the tf32 planes and still be exact
E2 = E[:200_000].clone(); E2[12345, 3] = 1.0e6
idx2 = CatalogIndex(w, E2)
os.environ["CFL_SCORE_LB_F16"] = "0"; r0 = idx2.rank_local(Pq[0], bench.TOPK)
os.environ["CFL_SCORE_LB_F16"] = "1"; r1 = idx2.rank_local(Pq[0], bench.TOPK)
print(json.dumps(dict(range_guard_results_equal=bool(torch.equal(r0[0], r1[0]) and torch.equal(r0[1], r1[1])))), flush=True)
