"""Achievable HBM bandwidth for different read:write mixes (torch ops, CUDA events)."""
import torch
dev = "cuda"
n = 128 * 1024 * 1024  # floats: 512 MB
a = torch.randn(n, device=dev); b = torch.empty(n, device=dev); c2 = torch.empty(2 * n, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timed(fn, reps=8):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[2:]); return ms[len(ms)//2]
for name, fn, byts in [
    ("copy 1R:1W", lambda: b.copy_(a), 8 * n),
    ("fill 0R:1W", lambda: b.zero_(), 4 * n),
    ("sum  1R:0W", lambda: a.sum(), 4 * n),
    ("cat  1R:2W", lambda: torch.cat((a, a), out=c2), 12 * n),
    ("add  2R:1W", lambda: torch.add(a, b, out=b), 12 * n),
    ("axpy 3R:1W", lambda: torch.addcmul(a, a, b, out=b), 16 * n),
]:
    t = timed(fn)
    print(f"{name}: {t*1e3:8.1f} us  {byts/t/1e6:7.0f} GB/s")
