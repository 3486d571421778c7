"""Fused train-step throughput (BASELINE config 2 shape: Dist model, F=4096, K=4, d=20) vs batch size,
with a per-kernel breakdown from CUDA events (run under gpurun)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compatibility-family-learning_b200"))
from cfl import _native as nat, ops, variables as vs
from cfl.models.dist import Dist

F, K, d = 4096, 4, 20
for B in (100, 4096, 65536):
    vs.reset_default_graph(); vs.set_seed(633)
    model = Dist(input_shape=(F,), latent_size=d, num_components=K, batch_size=B, lr=1e-3, beta1=0.9, beta2=0.999,
                 normalize_value=58.388599, data_normalizer=ops.normalizer(58.388599, 0.0))
    g = torch.Generator(device="cuda").manual_seed(B)
    batch = [torch.randn(B, F, generator=g, device="cuda").clamp_(min=0).mul_(20).clamp_(max=58.388599) for _ in range(4)]
    for _ in range(3): model.train_step(*batch)
    torch.cuda.synchronize()
    reps = 20 if B <= 4096 else 5
    t0 = time.time()
    for _ in range(reps): model.train_step(*batch)
    torch.cuda.synchronize()
    ms = (time.time() - t0) / reps * 1e3
    # kernel-level breakdown of the two heavy pieces
    def ev(fn, n=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    h = model.enc_src.proto
    y, _, _ = nat.project_fwd(batch[0], h.V, None, h.b, False, 1 / 58.388599, None)
    dy = torch.randn_like(y)
    fwd = ev(lambda: nat.project_fwd(batch[0], h.V, None, h.b, False, 1 / 58.388599, None))
    bwd = ev(lambda: nat.project_bwd(batch[0], h.V, None, h.b, False, 1 / 58.388599, None, None, None, dy))
    print(json.dumps(dict(B=B, step_ms=round(ms, 3), pairs_per_s=round(2 * B / ms * 1e3), proj_fwd_ms=round(fwd, 4),
                          proj_bwd_ms=round(bwd, 4), proj_fwd_tflops=round(2.0 * B * F * K * d / fwd / 1e9, 1),
                          proj_bwd_tflops=round(2.0 * B * F * K * d / bwd / 1e9, 1))), flush=True)


def bench_graph(B, F, K, d, reps=200):
    """Same model, step replayed as one CUDA graph (no host round trip)."""
    import time
    from cfl import variables as vs
    from cfl.models.dist import Dist
    from cfl.ops import normalizer
    vs.reset_default_graph()
    model = Dist(input_shape=(F,), latent_size=d, num_components=K, batch_size=B, lr=1e-3, beta1=0.9, beta2=0.999,
                 normalize_value=58.388599, data_normalizer=normalizer(58.388599, 0.0))
    g = torch.Generator(device="cuda").manual_seed(0)
    batch = [torch.randn(B, F, generator=g, device="cuda").clamp_(min=0) * 20 for _ in range(4)]
    for _ in range(5): model.train_step_graph(*batch)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): model.train_step_graph(*batch)
    torch.cuda.synchronize()
    ms = (time.time() - t0) / reps * 1e3
    print(json.dumps(dict(mode="cuda_graph", B=B, F=F, K=K, d=d, ms_per_step=round(ms, 4), pairs_per_s=round(2 * B / ms * 1e3))))


if __name__ == "__main__":
    for B in (100, 1000, 8192):
        bench_graph(B, 4096, 4, 20)
