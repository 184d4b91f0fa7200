"""Who launches the full-tensor copies of a training step: torch.profiler with Python stacks, grouped for the
aten::copy_ / add_ / add / slice_backward calls above a size threshold - development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from torch.profiler import profile, ProfilerActivity
from grafp_b200 import synth
from grafp_b200.encoder.graph_encoder import GraphEncoder
from grafp_b200.simclr.simclr import SimCLR
from grafp_b200.simclr.ntxent import ntxent_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = dict(synth.DEFAULT_CFG)
dev = torch.device("cuda")
model = SimCLR(cfg, GraphEncoder(cfg=cfg, in_channels=8, k=3)).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=8e-5)
s_i, s_j = (t.to(dev) for t in synth.synth_spec(B, 1))


def step():
    opt.zero_grad(set_to_none=True)
    _, _, z_i, z_j = model(s_i, s_j)
    loss = ntxent_loss(z_i.float(), z_j.float(), cfg)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
groups = collections.defaultdict(lambda: [0, 0.0])
SKIP = ("aten::convolution", "aten::_convolution", "aten::conv2d", "aten::cudnn_convolution", "aten::convolution_backward")
for ev in prof.events():
    if ev.name.startswith("aten::") and ev.name not in SKIP and ev.self_device_time_total > 0:
        stack = [s for s in (ev.stack or []) if "grafp_b200" in s or "scripts/" in s or "autograd" in s][:4]
        key = (ev.name, str(ev.input_shapes)[:80], " <- ".join(s.split("/")[-1][:60] for s in stack))
        groups[key][0] += 1
        groups[key][1] += ev.self_device_time_total
for key, (n, us) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us/1e3:8.2f} ms {n:4d}x {key[0]:22s} {key[1]:80s} {key[2]}")
