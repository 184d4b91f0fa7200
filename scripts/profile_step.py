"""Print the CUDA-time breakdown of one training step (torch.profiler) - development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from grafp_b200 import synth
from grafp_b200.encoder.graph_encoder import GraphEncoder
from grafp_b200.simclr.simclr import SimCLR
from grafp_b200.simclr.ntxent import ntxent_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
BF16 = "--bf16" in sys.argv
cfg = dict(synth.DEFAULT_CFG)
dev = torch.device("cuda")
model = SimCLR(cfg, GraphEncoder(cfg=cfg, in_channels=8, k=3)).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=8e-5)
s_i, s_j = (t.to(dev) for t in synth.synth_spec(B, 1))

def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=BF16):
        _, _, z_i, z_j = model(s_i, s_j)
    loss = ntxent_loss(z_i.float(), z_j.float(), cfg)
    loss.backward()
    opt.step()

for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
