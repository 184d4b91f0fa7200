"""Time the hot-path kernels alone at the benchmark batch (B = 512 segments): CUDA events on the launching stream,
L2 flushed between launches, median of 10.  fp32 and bf16, with the A/B options of include/grafp_b200.h.

    python scripts/bench_kernels.py [k1] [k23] [k5] [gemm] [--dtype fp32|bf16|both]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from grafp_b200 import ops  # noqa: E402

dev = "cuda"
B = int(os.environ.get("B", "512"))
k = int(os.environ.get("K", "3"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except OSError:
    PEAK = 6456.0
STAGES = [(1024, 64), (512, 128), (256, 256), (128, 512)]


def timed(fn, reps=10):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()                     # evict L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[2:])
    return ms[len(ms) // 2]


def rows(N, C, dtype, relu=True, grad=False):
    x = torch.randn(B, C, N, 1, device=dev)
    if relu:
        x = torch.relu(x)
    x = x.to(dtype).contiguous(memory_format=torch.channels_last)
    return x.requires_grad_(True) if grad else x


def correlated_rows(N, C, dtype):
    """Node features whose neighbouring rows are similar (a smooth random walk along the node axis) - closer to the
    encoder's feature maps than independent rows."""
    steps = 0.15 * torch.randn(B, N, C, device=dev)
    x = torch.cumsum(steps, dim=1) + torch.randn(B, 1, C, device=dev)
    return x.permute(0, 2, 1).unsqueeze(-1).to(dtype).contiguous(memory_format=torch.channels_last)


def bench_k1(dtype):
    print(f"=== K1 k-NN graph, dtype={dtype}, B={B}, k={k}: op = normalise launch + Gram/top-k launch")
    for name, make in (("independent rows", lambda N, C: rows(N, C, dtype, relu=False)), ("correlated rows", lambda N, C: correlated_rows(N, C, dtype))):
        for epi, label in ((0, "auto"), (3, "group-max"), (1, "vote"), (2, "queue")):
            ops.set_option("knn_epilogue", epi)
            line = f"{name:17s} epilogue={label:9s}"
            for (N, C) in STAGES:
                x = make(N, C)
                t = timed(lambda: ops.knn_graph(x, k))
                fl = 2.0 * B * N * N * C
                line += f" | N={N} C={C}: {t*1e3:6.1f} us {fl/t/1e9:6.1f} TF/s"
            print(line, flush=True)
    ops.set_option("knn_epilogue", 0)


def bench_k23(dtype):
    e = 4 if dtype == torch.float32 else 2
    print(f"=== K2 / K3, dtype={dtype}, B={B}, k={k}")
    for bwd_form, label in ((2, "cluster, bulk-staged"), (1, "cluster + device fence"), (3, "gather (deterministic)"), (0, "dense + scatter pair")):
        ops.set_option("mr_bwd_form", bwd_form)
        for (N, C) in STAGES:
            x = rows(N, C, dtype, grad=True)
            nbr, nbr32 = ops.knn_graph(x, k)
            out = ops.mr_aggregate(x, nbr32)
            t_fwd = timed(lambda: ops.mr_aggregate(x, nbr32))
            g = torch.randn_like(out)
            t_bwd = timed(lambda: torch.autograd.grad(out, x, g, retain_graph=True))
            b_fwd = B * (N * C * e + N * k * 4 + 2 * N * C * e + N * C)
            b_bwd = B * (2 * N * C * e + N * C + N * k * 4 + N * C * e)
            print(f"bwd={label:22s} N={N:5d} C={C:4d}  K2 {t_fwd*1e3:6.1f} us {b_fwd/t_fwd/1e6:6.0f} GB/s ({b_fwd/t_fwd/1e6/PEAK*100:4.1f}%)   "
                  f"K3 {t_bwd*1e3:6.1f} us {b_bwd/t_bwd/1e6:6.0f} GB/s ({b_bwd/t_bwd/1e6/PEAK*100:4.1f}%)", flush=True)
    ops.set_option("mr_bwd_form", 2)


def bench_k5(dtype):
    e = 4 if dtype == torch.float32 else 2
    print(f"=== K5 fused BatchNorm, dtype={dtype}, B={B}")
    for (N, C, mode) in [(1024, 64, "res"), (1024, 128, "relu"), (1024, 256, "relu"), (256, 256, "res"), (256, 1024, "relu"),
                         (128, 2048, "relu"), (128, 512, "plain")]:
        x = rows(N, C, dtype, relu=False, grad=True)
        res = torch.randn_like(x).requires_grad_(True)
        up = torch.randn_like(x)
        bn = torch.nn.BatchNorm2d(C).to(dev).train()
        S = x.numel() * e

        def ours():
            return ops.batch_norm_act(x, bn, relu=(mode == "relu"), residual=res if mode == "res" else None)

        def eager():
            y = bn(x)
            if mode == "res":
                y = y + res
            return torch.relu(y) if mode == "relu" else y

        line = f"N={N:5d} C={C:5d} {mode:5s} ({S/1e6:6.0f} MB/tensor)"
        for name, f, keep in (("ours", ours, 80), ("ours, no L2 hints", ours, 0), ("torch", eager, 80)):
            ops.set_option("bn_l2_keep_mb", keep)
            out = f()
            t_f = timed(f)
            ins = (x, res) if mode == "res" else (x,)
            t_b = timed(lambda: torch.autograd.grad(out, ins + tuple(bn.parameters()), up, retain_graph=True))
            passes_f = 3 + (1 if mode == "res" else 0)
            line += (f" | {name}: fwd {t_f*1e3:7.1f} us ({passes_f*S/t_f/1e6:5.0f} GB/s, {passes_f*S/t_f/1e6/PEAK*100:3.0f}%) "
                     f"bwd {t_b*1e3:7.1f} us ({5*S/t_b/1e6:5.0f} GB/s, {5*S/t_b/1e6/PEAK*100:3.0f}%)")
        ops.set_option("bn_l2_keep_mb", 80)
        print(line, flush=True)


def bench_gemm(dtype):
    """1x1 convolution + train-mode BatchNorm forward: tcgen05 GEMM with the statistics in its epilogue + apply pass
    (conv_gemm = 1) against cuDNN convolution + the two-pass BatchNorm kernel (conv_gemm = 0), TF32 allowed in both."""
    e = 4 if dtype == torch.float32 else 2
    torch.backends.cudnn.allow_tf32 = True
    print(f"=== 1x1 convolution (+ BatchNorm forward), dtype={dtype}, B={B}: GB/s over R (Cin + Cout) e")
    lib = ops._native.load()
    for (N, C) in STAGES:
        for (cin, cout, G) in ((C, C, 1), (2 * C, 2 * C, 4), (2 * C, C, 1), (C, 4 * C, 1), (4 * C, C, 1)):
            x = rows(N, cin, dtype, relu=True)
            conv = torch.nn.Conv2d(cin, cout, 1, bias=False, groups=G).to(dev)
            bn = torch.nn.BatchNorm2d(cout).to(dev).train()
            w = conv.weight.detach().to(dtype)
            bytes_conv = B * N * (cin + cout) * e
            t_own = timed(lambda: ops._conv1x1_stats_call(lib, x, w, G))
            if G > 1 and dtype == torch.bfloat16:   # (cuDNN's bf16 grouped kernel is the 29 ms one: its dense form instead)
                wd = ops._block_diag_weight(w, G)
                t_lib = timed(lambda: torch.nn.functional.conv2d(x, wd))
            else:
                t_lib = timed(lambda: torch.nn.functional.conv2d(x, w, groups=G))
            line = (f"N={N:5d} {cin:5d}->{cout:5d}{' g4' if G > 1 else '   '} | conv: ours {t_own*1e3:7.1f} us ({bytes_conv/t_own/1e6:5.0f} GB/s, "
                    f"{bytes_conv/t_own/1e6/PEAK*100:3.0f}%)  cuDNN {t_lib*1e3:7.1f} us ({bytes_conv/t_lib/1e6:5.0f} GB/s)")
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
                for flag in (2, 0):
                    ops.set_option("conv_gemm", flag)
                    xg = x.detach().requires_grad_(True)
                    t = timed(lambda: ops.conv_batch_norm_act(xg, conv, bn, relu=True))
                    line += f" | conv+BN+ReLU conv_gemm={flag}: {t*1e3:7.1f} us"
            ops.set_option("conv_gemm", 1)
            print(line, flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    which = set(args) or {"k1", "k23", "k5"}
    dt = "both"
    if "--dtype" in sys.argv:
        dt = sys.argv[sys.argv.index("--dtype") + 1]
    dtypes = {"fp32": [torch.float32], "bf16": [torch.bfloat16], "both": [torch.float32, torch.bfloat16]}[dt]
    print(f"HBM peak {PEAK:.0f} GB/s ({torch.cuda.get_device_name(0)})")
    for dtype in dtypes:
        if "k1" in which:
            bench_k1(dtype)
        if "k23" in which:
            bench_k23(dtype)
        if "k5" in which:
            bench_k5(dtype)
        if "gemm" in which:
            bench_gemm(dtype)
