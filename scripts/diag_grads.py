"""Diagnostic: per-parameter gradient error of the device encoder vs the oracle (CPU and same-GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from grafp_b200 import synth
from grafp_b200.encoder.graph_encoder import GraphEncoder
from grafp_b200.encoder.gcn_lib import torch_edge
from oracle import grafp_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
if len(sys.argv) > 1:
    torch.backends.fp32_precision = "ieee"
    torch.backends.cudnn.fp32_precision = "ieee"
    torch.backends.cudnn.conv.fp32_precision = "ieee"
    torch.backends.cuda.matmul.fp32_precision = "ieee"
print("fp32_precision:", getattr(torch.backends.cudnn.conv, "fp32_precision", None), getattr(torch.backends.cudnn, "fp32_precision", None))
DEV = "cuda"
cfg = dict(synth.DEFAULT_CFG)
enc = GraphEncoder(cfg=cfg, in_channels=8, k=3)
sd = enc.state_dict(); keep = {k: v for k, v in sd.items() if k.endswith("relative_pos")}
enc.load_state_dict(synth.synth_state_dict({k: v.shape for k, v in sd.items()}, 555, keep))
base = {k: v.clone() for k, v in enc.state_dict().items() if not k.endswith("relative_pos")}
trainable = [n for n, q in enc.named_parameters() if q.requires_grad]
g = torch.Generator().manual_seed(8)
x = torch.rand(4, 8, 1024, generator=g); up = torch.randn(4, 1024, generator=g)
enc.to(DEV).train()
rec = []
for m in enc.modules():
    if isinstance(m, torch_edge.DenseDilatedKnnGraph):
        m.register_forward_hook(lambda mod, inp, out: rec.append(out[0].detach().cpu()))
xg = x.to(DEV).requires_grad_(True)
out = enc(xg); (out * up.to(DEV)).sum().backward()

def run_oracle(device):
    p = {k: v.clone().to(device) for k, v in base.items()}
    for n in trainable: p[n].requires_grad_(True)
    xo = x.clone().to(device).requires_grad_(True)
    ref = O.graph_encoder(p, xo, True, k=3, graph_fn=O.GraphReplay(rec))
    (ref * up.to(device)).sum().backward()
    return p, xo, ref

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

grads = dict(enc.named_parameters())
for device in ("cpu", "cuda"):
    p, xo, ref = run_oracle(device)
    print(f"== oracle on {device}: out {rel(out, ref):.2e}  x.grad {rel(xg.grad, xo.grad):.2e}")
    for n in trainable:
        if n.endswith("weight") and grads[n].dim() == 4:
            print(f"   {n:55s} {rel(grads[n].grad, p[n].grad):.2e}")
    if device == "cpu":
        pc, xc = p, xo
    else:
        print(f"== oracle cuda vs oracle cpu: x.grad {rel(xo.grad, xc.grad):.2e}")
        for n in trainable:
            if n.endswith("weight") and p[n].dim() == 4 and ("fc1.0" in n or "stem" in n):
                print(f"   {n:55s} {rel(p[n].grad, pc[n].grad):.2e}")
