"""Fused train-mode BatchNorm (+ReLU / +residual) against the PyTorch op sequence it replaces, at the encoder's
shapes (B = 512): CUDA events, L2 flushed between launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grafp_b200 import ops

dev = "cuda"
B = int(os.environ.get("B", "512"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
PEAK = 6550.0


def timed(fn, reps=8):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[2:])
    return ms[len(ms) // 2]


for (N, C, mode) in [(1024, 64, "res"), (1024, 128, "relu"), (1024, 256, "relu"), (256, 256, "res"), (256, 1024, "relu"),
                     (128, 2048, "relu"), (128, 512, "plain")]:
    x = torch.randn(B, C, N, 1, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    res = torch.randn_like(x).requires_grad_(True)
    up = torch.randn_like(x)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    S = x.numel() * 4

    def ours():
        return ops.batch_norm_act(x, bn, relu=(mode == "relu"), residual=res if mode == "res" else None)

    def eager():
        y = bn(x)
        if mode == "res":
            y = y + res
        return torch.relu(y) if mode == "relu" else y

    line = f"N={N:5d} C={C:5d} {mode:5s} ({S/1e6:6.0f} MB/tensor)"
    for name, f in (("ours", ours), ("torch", eager)):
        out = f()
        t_f = timed(f)
        ins = (x, res) if mode == "res" else (x,)
        t_b = timed(lambda: torch.autograd.grad(out, ins + tuple(bn.parameters()), up, retain_graph=True))
        passes_f = 3 + (1 if mode == "res" else 0)
        line += f" | {name}: fwd {t_f*1e3:7.1f} us ({passes_f*S/t_f/1e6:5.0f} GB/s of {passes_f}S) bwd {t_b*1e3:7.1f} us ({5*S/t_b/1e6:5.0f} GB/s of 5S)"
    print(line, flush=True)
