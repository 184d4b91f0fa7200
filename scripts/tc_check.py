"""Development check: the tcgen05 k-NN kernels (f16x3 and tf32x3) vs the exact-fp32 SIMT path on the
device, then timings at the benchmark shapes (run under `timeout`)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grafp_b200 import _native, ops

dev = "cuda"
cases = [  # B, C, N, k, d
    (2, 64, 256, 3, 1), (2, 64, 1024, 3, 1), (2, 128, 512, 3, 1), (2, 256, 256, 3, 1), (2, 512, 128, 3, 1),
    (2, 64, 1024, 8, 1), (2, 64, 1024, 4, 2), (3, 72, 300, 5, 1), (2, 256, 1024, 3, 1), (2, 256, 200, 3, 1),
    (2, 64, 1024, 16, 1), (64, 64, 1024, 3, 1), (64, 256, 256, 3, 1),
]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for (B, C, N, k, d) in cases:
    g = torch.Generator(device=dev).manual_seed(B * 1000 + N + C)
    x = torch.randn(B, C, N, 1, device=dev, generator=g)
    ref, _ = ops.knn_graph(x, k, d, algo=_native.KNN_SIMT)
    torch.cuda.synchronize()
    for algo in (_native.KNN_TC, _native.KNN_TC_TF32):
        t0 = time.time()
        got, _ = ops.knn_graph(x, k, d, algo=algo)
        torch.cuda.synchronize()
        agree = float((got == ref).float().mean())
        rank0 = float((got[..., 0] == torch.arange(N, device=dev)).float().mean())
        print(f"B={B} C={C} N={N} k={k} d={d}: {ops.knn_last_variant():7s} agree={agree:.6f} rank0_self={rank0:.4f} "
              f"({(time.time()-t0)*1e3:.1f} ms)", flush=True)
        if agree < 0.999:
            bad = (got != ref).nonzero()[:5]
            for bb, nn_, jj in bad.tolist():
                print("   mismatch at", bb, nn_, jj, "tc", got[bb, nn_].tolist(), "simt", ref[bb, nn_].tolist())
# timing at the benchmark shape
for (B, C, N) in [(512, 64, 1024), (512, 128, 512), (512, 256, 256), (512, 512, 128)]:
    x = torch.randn(B, C, N, 1, device=dev)
    for algo, name in ((_native.KNN_TC_TF32, "tf32x3"), (_native.KNN_TC, "f16x3")):
        for _ in range(2):
            ops.knn_graph(x, 3, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.knn_graph(x, 3, algo=algo)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"B={B} C={C} N={N} {name}: {ms:.3f} ms  {2.0*B*N*N*C/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)
