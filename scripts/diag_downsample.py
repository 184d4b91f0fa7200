"""Downsample: 3-tap pointwise form against the module's own 3x3 stride-2 convolution (fp32, TF32 off), and both
against fp64 - how much rounding noise does each carry?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grafp_b200 import ops
from grafp_b200.encoder.graph_encoder import Downsample
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
for (B, C, N) in [(32, 64, 1024), (32, 128, 512), (32, 256, 256)]:
    m = Downsample(C, 2 * C).cuda().train()
    x = torch.relu(torch.randn(B, C, N, 1, device="cuda")).contiguous(memory_format=torch.channels_last)
    m64 = Downsample(C, 2 * C).double().cuda().train()
    m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in m.state_dict().items()})
    ref64 = m64.conv(x.double())
    for grad in (True, False):
        with torch.set_grad_enabled(grad):
            os.environ["GRAFP_FUSED_BN"] = "1"
            a = m(x)
            os.environ["GRAFP_FUSED_BN"] = "0"
            b = m(x)
        ea = float((a.double() - ref64).norm() / ref64.norm()); eb = float((b.double() - ref64).norm() / ref64.norm())
        print(f"B={B} C={C} N={N} grad={grad}: 3-tap form err {ea:.2e}   cuDNN 3x3 err {eb:.2e}   max|a-b| {float((a-b).abs().max()):.2e}")
