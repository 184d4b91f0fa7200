"""Which big elementwise adds / copies does one training step issue, and with which strides?  (The ncu launch list
shows ~48 large at::elementwise_kernel<128, 2, add> launches per step - the strided, non-vectorised TensorIterator
path - next to the vectorised ones.)  Logs every aten add / copy over >= 1M elements with its operand strides."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
from grafp_b200 import synth
from grafp_b200.encoder.graph_encoder import GraphEncoder
from grafp_b200.simclr.simclr import SimCLR
from grafp_b200.simclr.ntxent import ntxent_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seen = collections.Counter()


class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if ("add" in name or "copy_" in name) and "addmm" not in name:
            ts = [a for a in args if isinstance(a, torch.Tensor)]
            if ts and max(t.numel() for t in ts) >= (1 << 20):
                key = (name, tuple((tuple(t.shape), tuple(t.stride())) for t in ts))
                seen[key] += 1
        return func(*args, **(kwargs or {}))


cfg = dict(synth.DEFAULT_CFG)
dev = torch.device("cuda")
model = SimCLR(cfg, GraphEncoder(cfg=cfg, in_channels=8, k=3)).to(dev).train()
s_i, s_j = (t.to(dev) for t in synth.synth_spec(B, 1))
for it in range(2):
    model.zero_grad(set_to_none=True)
    ctx = Log() if it == 1 else torch.no_grad.__new__(torch.no_grad)  # log the second step only
    if it == 1:
        with ctx:
            _, _, z_i, z_j = model(s_i, s_j)
            loss = ntxent_loss(z_i, z_j, cfg)
            loss.backward()
    else:
        _, _, z_i, z_j = model(s_i, s_j)
        ntxent_loss(z_i, z_j, cfg).backward()
torch.cuda.synchronize()
for (name, sig), n in sorted(seen.items(), key=lambda kv: -kv[1]):
    print(n, name, " | ".join(f"{s}:{st}" for s, st in sig))
