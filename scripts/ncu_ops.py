"""Run the hot-path kernels alone at the benchmark batch (for `ncu --set full` captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grafp_b200 import ops

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dtype = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
for (N, C) in [(1024, 64), (512, 128), (256, 256), (128, 512)]:
    x = torch.relu(torch.randn(B, C, N, 1, device=dev)).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(2 * C).to(dev).train()
    for _ in range(reps):
        nbr, nbr32 = ops.knn_graph(x, 3)
        out = ops.mr_aggregate(x, nbr32)
        y = ops.batch_norm_act(out, bn, relu=True)           # fused train-mode BatchNorm + ReLU on the (B, N, 2C) rows
        g = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, g)
    # FFN fc1 of the stage (C -> 4C) through the tcgen05 convolution with the statistics epilogue + the apply pass
    torch.backends.cudnn.allow_tf32 = True
    conv = torch.nn.Conv2d(C, 4 * C, 1, bias=False).to(dev)
    bn4 = torch.nn.BatchNorm2d(4 * C).to(dev).train()
    gconv = torch.nn.Conv2d(2 * C, 2 * C, 1, groups=4).to(dev)   # BasicConv of the graph convolution on the K2 output
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        for _ in range(reps):
            ops.set_option("conv_gemm", 2)
            h = ops.conv_batch_norm_act(x, conv, bn4, relu=True)
            hg = ops.conv_batch_norm_act(out.detach(), gconv, bn, relu=True)
    ops.set_option("conv_gemm", 1)
    if N >= 256:
        taps = ops._DownsampleTaps.apply(x)
        taps.backward(torch.ones_like(taps))
    torch.cuda.synchronize()
# the 8f kernels at the bench shapes: NT-Xent on 1024 x 128 embeddings, the peak extractor on 512 segments
from grafp_b200 import synth
from grafp_b200.peak_extractor import GPUPeakExtractorv2
z = torch.nn.functional.normalize(torch.randn(1024, 128, device=dev), dim=1).requires_grad_(True)
loss = ops.ntxent(z, 0.05)
loss.backward()
pe = GPUPeakExtractorv2(dict(synth.DEFAULT_CFG)).to(dev)
spec = synth.synth_spec(B, 3)[0].to(dev)
pts = pe(spec)
pts.sum().backward()
torch.cuda.synchronize()
print("done", ops.knn_last_algo())
