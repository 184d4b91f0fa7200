import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grafp_b200 import synth, ops
from grafp_b200.encoder.graph_encoder import GraphEncoder
from grafp_b200.simclr.simclr import SimCLR
from grafp_b200.simclr import ntxent as ntx
from grafp_b200.training import GraphedTrainStep

cfg = dict(synth.DEFAULT_CFG)
DEV = "cuda"
B = 6
def build():
    torch.manual_seed(0)
    m = SimCLR(cfg, GraphEncoder(cfg=cfg, in_channels=8, k=3)).to(DEV).train()
    return m, torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True)
def loss_of(h_i, h_j, z_i, z_j):
    return ntx.ntxent_loss(z_i.float(), z_j.float(), cfg)
batches = [tuple(t.to(DEV) for t in synth.synth_spec(B, 900 + i)) for i in range(4)]

def run(tag):
    m_g, opt_g = build()
    m_e, opt_e = build()
    gstep = GraphedTrainStep(m_g, opt_g, loss_of, list(batches[0]), warmup=1)
    torch.cuda.synchronize()
    m_e.load_state_dict(m_g.state_dict()); opt_e.load_state_dict(opt_g.state_dict())
    lg = float(gstep(*batches[1]))
    opt_e.zero_grad(set_to_none=True)
    le = loss_of(*m_e(*batches[1])); le.backward(); opt_e.step()
    diffs = []
    for (n, a), (_, b) in zip(m_g.named_parameters(), m_e.named_parameters()):
        if a.grad is not None and b.grad is not None:
            diffs.append((float((a.grad - b.grad).norm() / (b.grad.norm() + 1e-20)), n, float(b.grad.norm())))
    diffs.sort(reverse=True)
    print(f"[{tag}] loss graph {lg:.6f} eager {float(le):.6f}; worst grad diffs:", [(round(d, 4), n, f"{g:.2e}") for d, n, g in diffs[:6]], "n>1e-3:", sum(d > 1e-3 for d, _, _ in diffs), "of", len(diffs), flush=True)

run("default")
ops.set_option("bn_persistent", 0); run("bn two-kernel"); ops.set_option("bn_persistent", 1)
ops.set_option("mr_bwd_form", 0); run("K3 pair"); ops.set_option("mr_bwd_form", 2)
ops.peak_extract_supported = lambda *a, **k: False; run("no fused peak")
ops.ntxent_supported = lambda *a, **k: False; run("no fused peak, no fused ntxent")
