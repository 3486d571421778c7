"""Host time to ENQUEUE one ranking step (no back-pressure: 20 steps on an idle stream) vs its GPU time, 1 rank or torchrun."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
import bench
from cfl import _native as nat
from cfl.ranking import CatalogIndex

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
w = bench.synth_weights(dev)
E = torch.randn(bench.N_PER_GPU, bench.D, device=dev) * 0.2
index = CatalogIndex(w, E, idx_base=rank * bench.N_PER_GPU, n_total=bench.N_PER_GPU * world)
xq = bench.synth_features(bench.Q, dev, bench.SEED + 7)
for _ in range(5):
    r = index.rank_async(xq, bench.TOPK)
torch.cuda.synchronize()
if world > 1: torch.distributed.barrier()
res = []
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    outs = [index.rank_async(xq, bench.TOPK) for _ in range(20)]
    t1 = time.perf_counter(); e1.record()
    for o in outs: o[2].synchronize()
    torch.cuda.synchronize()
    res.append(((t1 - t0) * 1e3 / 20, e0.elapsed_time(e1) / 20))
if rank == 0:
    print(json.dumps(dict(world=world, host_enqueue_ms_per_step=[round(a, 3) for a, _ in res], gpu_ms_per_step=[round(b, 3) for _, b in res])))
if world > 1:
    torch.distributed.destroy_process_group()
