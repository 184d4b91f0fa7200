"""Self-measured roofline denominators on this box (same recipe as the driver's MEASURED_PEAKS.json:
a STREAM-style device copy and a cuBLAS bf16 GEMM, burst and sustained).  Prints one JSON line."""
import json, time
import torch

dev = "cuda"
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    b.copy_(a)
e1.record(); torch.cuda.synchronize()
hbm = 2 * n * 20 / (e0.elapsed_time(e1) / 1e3) / 1e9
del a, b
m = 8192
x = torch.randn(m, m, device=dev, dtype=torch.bfloat16)
y = torch.randn(m, m, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    x @ y
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    x @ y
e1.record(); torch.cuda.synchronize()
burst = 2 * m ** 3 * 10 / (e0.elapsed_time(e1) / 1e3) / 1e12
t0 = time.time(); it = 0
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(20):
        x @ y
    it += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * m ** 3 * it / (e0.elapsed_time(e1) / 1e3) / 1e12
print(json.dumps({"hbm_gbs": hbm, "bf16_tflops": burst, "bf16_tflops_sustained": sust, "source": "scripts/measure_peaks.py"}))
