"""Per-phase device time of the sharded ranking step at world > 1 (torchrun)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
import bench as B
from cfl import _native as nat
from cfl.ranking import CatalogIndex

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
torch.distributed.init_process_group("nccl", device_id=dev)
w = B.synth_weights(dev)
N = 1_000_000
E = torch.empty(N, B.D, device=dev)
for lo in range(0, N, 1 << 18):
    hi = min(N, lo + (1 << 18))
    E[lo:hi] = nat.project_fwd(B.synth_features(hi - lo, dev, B.SEED + 100 + rank * 64 + lo // (1 << 18)), w.V0, w.g0, w.b0, True, w.in_scale, None)[0]
index = CatalogIndex(w, E, idx_base=rank * N, n_total=N * world)
xq = B.synth_features(B.Q, dev, B.SEED + 7)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(8):
    index.rank(xq, 100)
torch.cuda.synchronize(); torch.distributed.barrier()
T = {"local": 0.0, "gather": 0.0, "merge": 0.0}
steps = 40
t0 = time.time()
for it in range(steps):
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record()
    Pq = index.project_queries(xq); tv, ti = index.rank_local(Pq, 100)
    b.record()
    gv = torch.empty(world * B.Q, 100, dtype=tv.dtype, device=dev); gi = torch.empty(world * B.Q, 100, dtype=ti.dtype, device=dev)
    torch.distributed.all_gather_into_tensor(gv, tv); torch.distributed.all_gather_into_tensor(gi, ti)
    c.record()
    nat.topk_merge(gv.view(world, B.Q, 100), gi.view(world, B.Q, 100))
    d.record()
    if it % 10 == 9:
        torch.cuda.synchronize()
        T["local"] += a.elapsed_time(b); T["gather"] += b.elapsed_time(c); T["merge"] += c.elapsed_time(d)
cpu_issue = (time.time() - t0) / steps * 1e3
torch.cuda.synchronize()
wall = (time.time() - t0) / steps * 1e3
index.rank(xq, 100); torch.cuda.synchronize()
print(f"rank {rank}: sampled steps local {T['local']/4:.3f} ms gather {T['gather']/4:.3f} ms merge {T['merge']/4:.3f} ms | cpu issue {cpu_issue:.3f} ms/step, wall {wall:.3f} ms/step", flush=True)
torch.distributed.destroy_process_group()
