"""TF32 dense tensor-core peak on this B200 ('the builder must add a TF32 peak probe'; TF32 is not in
MEASURED_PEAKS.json).  cuBLAS fp32 matmul with TF32 allowed, 8192^3: best of 10 (burst) and back to back for 3 s
(sustained), same method as the driver's bf16 probe.  Writes gpurun_out/tf32_peak.json."""
import json, os, time
import torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); a @ b; e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0, reps = time.time(), 0
s.record()
while time.time() - t0 < 3.0:
    for _ in range(20): a @ b
    reps += 20
    torch.cuda.synchronize()
e.record(); torch.cuda.synchronize()
fl = 2.0 * n ** 3
out = dict(tf32_tflops=round(fl / best / 1e9, 1), tf32_tflops_sustained=round(fl * reps / s.elapsed_time(e) / 1e9, 1),
           how="torch.matmul fp32 with allow_tf32, 8192^3, best of 10 / back to back 3 s", gpu=torch.cuda.get_device_name(0))
# bf16 beside it on the same box, same method (to relate to MEASURED_PEAKS.json)
ab, bb = a.bfloat16(), b.bfloat16()
for _ in range(3): ab @ bb
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ab @ bb; e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
out["bf16_tflops_same_box"] = round(fl / best / 1e9, 1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tf32_peak.json", "w"), indent=1)
print(json.dumps(out))
