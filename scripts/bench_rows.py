"""Measured numbers for the SURVEY 8 rows that bench.py's training step does not exercise (one B200, CUDA events,
L2 flushed between launches):

  * the EdgeConv / plain-gather kernels (K2e, K3e, K4, batched_index_select) at the encoder's stage shapes,
    as achieved HBM GB/s over the algorithmic bytes of SURVEY 8(d);
  * configs[3], the dense-graph stress: 1024 nodes, k = 16, dilation 1..4 (K = 16..64): K1, K2, K3 alone, next to
    the PyTorch-eager form of the same reference ops (normalize -> matmul -> topk; index + max) on the same GPU;
  * configs[4], fingerprint generation: the inference-only encoder over 10 000 synthetic segments in chunks of
    128 (generate.py:41) and 1024, segments/s.

The eager forms below are the reference's op sequence written inline (torch_edge.py:7-18,70-103,270-284,
torch_nn.py:79-98, torch_vertex.py:21-32); they are the "kernel to beat" timing, not a checker.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from grafp_b200 import ops, synth

dev = "cuda"
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6538.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=8):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[2:])
    return ms[len(ms) // 2]


def rows(B, C, N):
    return torch.relu(torch.randn(B, C, N, 1, device=dev)).contiguous(memory_format=torch.channels_last)


def gbs(nbytes, ms):
    return f"{ms*1e3:7.1f} us {nbytes/ms/1e6:6.0f} GB/s ({nbytes/ms/1e6/PEAK*100:4.1f}%)"


def eager_knn(x, k, d):
    with torch.no_grad():
        xn = F.normalize(x, p=2.0, dim=1).transpose(2, 1).squeeze(-1)
        inner = -2 * torch.matmul(xn, xn.transpose(2, 1))
        sq = torch.sum(xn * xn, dim=-1, keepdim=True)
        dist = sq + inner + sq.transpose(2, 1)
        _, nn_idx = torch.topk(-dist, k=k * d)
        return nn_idx[:, :, ::d]


def eager_mr(x, idx):
    B, C, N, _ = x.shape
    k = idx.shape[-1]
    base = torch.arange(0, B, device=x.device).view(-1, 1, 1) * N
    flat = (idx + base).contiguous().view(-1)
    xt = x.transpose(2, 1).contiguous().view(B * N, -1)
    x_j = xt[flat].view(B, N, k, C).permute(0, 3, 1, 2)
    m, _ = torch.max(x_j - x, -1, keepdim=True)
    return torch.cat([x.unsqueeze(2), m.unsqueeze(2)], dim=2).reshape(B, 2 * C, N, 1)


def edge_rows(B=512, k=3):
    print(f"--- EdgeConv / gather kernels at the encoder stages, B = {B}, k = {k}, fp32 (peak {PEAK:.0f} GB/s)")
    e = 4
    for (N, C) in [(1024, 64), (512, 128), (256, 256), (128, 512)]:
        x = rows(B, C, N).requires_grad_(True)
        _, n32 = ops.knn_graph(x, k)
        h = ops.edge_features(x, n32)                      # K2e: (B, 2C, N, k)
        t_e = timed(lambda: ops.edge_features(x, n32))
        g = torch.randn_like(h)
        t_eb = timed(lambda: torch.autograd.grad(h, x, g, retain_graph=True))   # K3e
        hd = h.detach().requires_grad_(True)
        mx = ops.max_over_k(hd)                            # K4
        t_m = timed(lambda: ops.max_over_k(hd))
        gm = torch.randn_like(mx)
        t_mb = timed(lambda: torch.autograd.grad(mx, hd, gm, retain_graph=True))
        xg = ops.gather_neighbors(x, n32)                  # batched_index_select
        t_g = timed(lambda: ops.gather_neighbors(x, n32))
        gg = torch.randn_like(xg)
        t_gb = timed(lambda: torch.autograd.grad(xg, x, gg, retain_graph=True))
        b_e = B * (N * C * e + N * k * 4 + 2 * k * N * C * e)
        b_eb = B * (2 * k * N * C * e + N * k * 4 + N * C * e)
        b_m = B * (k * N * 2 * C * e + N * 2 * C * e + N * 2 * C)
        b_mb = B * (N * 2 * C * e + N * 2 * C + k * N * 2 * C * e)
        b_g = B * (N * C * e + N * k * 4 + k * N * C * e)
        b_gb = B * (k * N * C * e + N * k * 4 + N * C * e)
        print(f"N={N:5d} C={C:4d}  K2e {gbs(b_e, t_e)}  K3e {gbs(b_eb, t_eb)}  K4 fwd {gbs(b_m, t_m)}  K4 bwd {gbs(b_mb, t_mb)}"
              f"  gather {gbs(b_g, t_g)}  gather bwd {gbs(b_gb, t_gb)}", flush=True)


def stress(B=256):
    print(f"--- configs[3] dense-graph stress: B = {B}, N = 1024, k = 16, dilation 1..4")
    e = 4
    N, k = 1024, 16
    for C in (64, 256):
        x = torch.randn(B, C, N, 1, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        xe = x.detach().contiguous()
        for d in (1, 2, 3, 4):
            nn64, n32 = ops.knn_graph(x, k, d)
            t_k = timed(lambda: ops.knn_graph(x, k, d), reps=5)
            var = ops.knn_last_variant()
            t_ke = timed(lambda: eager_knn(xe, k, d), reps=3)
            same = float((eager_knn(xe, k, d) == nn64).float().mean())
            line = f"C={C:4d} d={d}  K1[{var}] {t_k*1e3:8.1f} us ({2.0*B*N*N*C/t_k/1e9:6.1f} TF/s alg)  eager {t_ke*1e3:8.1f} us  x{t_ke/t_k:5.1f}  ids equal {same*100:.3f}%"
            if d == 1:
                out = ops.mr_aggregate(x, n32)
                t_f = timed(lambda: ops.mr_aggregate(x, n32), reps=5)
                g = torch.randn_like(out)
                t_b = timed(lambda: torch.autograd.grad(out, x, g, retain_graph=True), reps=5)
                xe2 = xe.clone().requires_grad_(True)
                oe = eager_mr(xe2, nn64)
                t_fe = timed(lambda: eager_mr(xe2, nn64), reps=3)
                ge = torch.randn_like(oe)
                t_be = timed(lambda: torch.autograd.grad(oe, xe2, ge, retain_graph=True), reps=3)
                b_f = B * (N * C * e + N * k * 4 + 2 * N * C * e + N * C)
                b_b = B * (2 * N * C * e + N * C + N * k * 4 + N * C * e)
                line += f"\n           K2 {gbs(b_f, t_f)} eager {t_fe*1e3:8.1f} us   K3 {gbs(b_b, t_b)} eager {t_be*1e3:8.1f} us"
                del out, oe, xe2
            print(line, flush=True)


def fingerprints(total=10000):
    from grafp_b200.encoder.graph_encoder import GraphEncoder
    from grafp_b200.peak_extractor import GPUPeakExtractorv2
    print(f"--- configs[4] fingerprint generation: inference-only encoder over {total} synthetic segments")
    cfg = dict(synth.DEFAULT_CFG)
    torch.manual_seed(0)
    enc = GraphEncoder(cfg=cfg, in_channels=cfg["n_filters"], k=3).to(dev).eval()
    pe = GPUPeakExtractorv2(cfg).to(dev).eval()
    spec = synth.synth_spec(1024, seed=7)[0].to(dev)
    for chunk in (128, 1024):
        xs = spec[:chunk]
        with torch.no_grad():
            for _ in range(2):
                enc(pe(xs))
            torch.cuda.synchronize()
            n_chunks = (total + chunk - 1) // chunk
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_chunks):
                fp = F.normalize(enc(pe(xs)), dim=1)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"chunk {chunk:5d}: {n_chunks * chunk / ms * 1e3:9.0f} segments/s ({ms / n_chunks:.2f} ms per chunk)", flush=True)
        from grafp_b200.inference import GraphedEncoder
        runner = GraphedEncoder(lambda s: F.normalize(enc(pe(s)), dim=1), xs)
        with torch.no_grad():
            same = torch.equal(runner(xs), F.normalize(enc(pe(xs)), dim=1))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n_chunks):
            fp = runner(xs)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"chunk {chunk:5d}: {n_chunks * chunk / ms * 1e3:9.0f} segments/s ({ms / n_chunks:.2f} ms per chunk)  CUDA-graph replay, "
              f"output identical to eager: {same}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["edge", "stress", "fp"]
    if "edge" in what:
        edge_rows()
    if "stress" in what:
        stress()
    if "stress1" in what:   # one launch set for ncu captures
        x = torch.randn(256, 64, 1024, 1, device=dev).contiguous(memory_format=torch.channels_last)
        ops.knn_graph(x, 16, 1)
        torch.cuda.synchronize()
    if "fp" in what:
        fingerprints()
