"""Time the hot-path kernels alone at the benchmark batch (CUDA events, L2 flushed between launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grafp_b200 import _native, ops

dev = "cuda"
B = int(os.environ.get("B", "512"))
k = int(os.environ.get("K", "3"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
PEAK = 6538.0


def timed(fn, reps=10):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()                     # evict L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[2:])
    return ms[len(ms) // 2]


def run(label):
    print(f"--- {label}: fwd={os.environ.get('GRAFP_MR_FWD_VARIANT','default')} bwd={os.environ.get('GRAFP_MR_BWD_VARIANT','default')}")
    for (N, C) in [(1024, 64), (512, 128), (256, 256), (128, 512)]:
        x = torch.relu(torch.randn(B, C, N, 1, device=dev)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        nbr, nbr32 = ops.knn_graph(x, k)
        t_knn = timed(lambda: ops.knn_graph(x, k))
        out = ops.mr_aggregate(x, nbr32)
        t_fwd = timed(lambda: ops.mr_aggregate(x, nbr32))
        g = torch.randn_like(out)
        t_bwd = timed(lambda: torch.autograd.grad(out, x, g, retain_graph=True))
        e = 4
        b_fwd = B * (N * C * e + N * k * 4 + 2 * N * C * e + N * C)
        b_bwd = B * (2 * N * C * e + N * C + N * k * 4 + N * C * e)
        fl = 2.0 * B * N * N * C
        print(f"N={N:5d} C={C:4d}  knn[{ops.knn_last_variant()}] {t_knn*1e3:6.1f} us ({fl/t_knn/1e9:6.1f} TF/s alg)   "
              f"mr_fwd {t_fwd*1e3:6.1f} us {b_fwd/t_fwd/1e6:6.0f} GB/s ({b_fwd/t_fwd/1e6/PEAK*100:4.1f}%)   "
              f"mr_bwd {t_bwd*1e3:6.1f} us {b_bwd/t_bwd/1e6:6.0f} GB/s ({b_bwd/t_bwd/1e6/PEAK*100:4.1f}%)", flush=True)


run("defaults (pipelined forward, TMA-staged cluster backward)")
if os.environ.get("ALL_VARIANTS", "1") == "1":
    os.environ["GRAFP_MR_FWD_VARIANT"] = "4"; os.environ["GRAFP_MR_BWD_VARIANT"] = "2"
    run("register-prefetch forward, cluster-fused atomic backward")
