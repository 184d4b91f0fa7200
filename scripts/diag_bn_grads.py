"""A/B of the encoder gradient accuracy (against the fp64 oracle) with the fused BatchNorm ops on and off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as T
from oracle import grafp_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(flag):
    os.environ["GRAFP_FUSED_BN"] = flag
    g = torch.Generator().manual_seed(8)
    x = torch.rand(6, 8, 1024, generator=g); up = torch.randn(6, 1024, generator=g)
    rows = []
    orig = T.assert_grads_as_accurate_as_reference
    def probe(ours, ref32, ref64, names, slack=3.0):
        scale = max(float(ref64[n].double().norm()) for n in names)
        for n in names:
            g64 = ref64[n].double(); denom = max(float(g64.norm()), 0.1 * scale)
            rows.append((float((ours[n].double().cpu() - g64).norm()) / denom, float((ref32[n].double() - g64).norm()) / denom, n))
        return 0.0
    T.assert_grads_as_accurate_as_reference = probe
    try:
        T._encoder_vs_oracle(555, x, up)
    finally:
        T.assert_grads_as_accurate_as_reference = orig
    rows.sort(key=lambda r: -(r[0] / (3 * r[1] + 1e-4)))
    print(f"GRAFP_FUSED_BN={flag}: worst e_ours/(3 e_ref + 1e-4):")
    for r in rows[:6]:
        print(f"   {r[0]:.2e} vs ref {r[1]:.2e}  ratio {r[0]/(3*r[1]+1e-4):.2f}  {r[2]}")
    print("   median e_ours/e_ref", sorted(r[0] / max(r[1], 1e-12) for r in rows)[len(rows) // 2])

run("0"); run("1")
