"""Sweeps the sampling knobs of cfl_score_topk_packed on the bench.py workload (C3: projected catalog of 1M rows,
Q=1024, top-100) inside one process: the library reads CFL_SCORE_* with getenv at every call (run under gpurun)."""
import itertools, json, os, sys
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
xq = [bench.synth_features(bench.Q, dev, bench.SEED + 7 + i) for i in range(4)]
Pq = [index.project_queries(x) for x in xq]


def run(reps=24):
    for i in range(3): index.rank_local(Pq[i % 4], bench.TOPK)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): index.rank_local(Pq[i % 4], bench.TOPK)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ref = [index.rank_local(p, bench.TOPK) for p in Pq]
base = run()
print(json.dumps(dict(knobs="default", ms=round(base, 4))), flush=True)
GRID = [tuple(int(v) for v in x.split(':')) for x in os.environ.get('CFL_KNOB_GRID', '').split(',') if x] or \
    list(itertools.product((12, 16, 24, 32), (3, 4, 6)))
for stride, mult in GRID:
    os.environ["CFL_SCORE_SAMPLE_STRIDE"] = str(stride)
    os.environ["CFL_SCORE_OPT_MULT"] = str(mult)
    same = all(torch.equal(index.rank_local(p, bench.TOPK)[1], r[1]) for p, r in zip(Pq, ref))
    print(json.dumps(dict(stride=stride, opt_mult=mult, ms=round(run(), 4), same_result=same)), flush=True)
