"""Small query batches against the C3 catalog: step / dominant-kernel time and filter statistics for the single adaptive
pass (CFL_SCORE_MIN_TILES=100000) and the two-pass path (=2), Q in {1, 4, 16, 32, 64, 128, 256}."""
import os, sys, torch
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
index = CatalogIndex(w, E)
for q in (1, 4, 16, 32, 64, 128, 256):
    for mt in ("100000", "2", None):
        if mt is None:
            os.environ.pop("CFL_SCORE_MIN_TILES", None)
        else:
            os.environ["CFL_SCORE_MIN_TILES"] = mt
        r = bench.small_q_line(index, dev, q=q, steps=30)
        st = index.rank_local_stats(bench.synth_features(q, dev, bench.SEED + 99), 100)[2]
        print("Q %4d min_tiles %-7s step %.4f ms  kernel %.4f ms  survivors/q %.0f redo %d" % (q, mt or "default", r["ms_per_step"], r["kernel_ms"], st["survivors"] / q, st["redo_queries"]), flush=True)
