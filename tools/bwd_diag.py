import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import synth
from helpers import cu
from adaptive_city_nerf_b200 import ops

def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))

sd = synth.make_expert_params(5, log2T=4)
wt = [cu(w) for w in synth.expert_weight_list(sd)]
for P, scale in ((128, 1e-7), (128, 1.0), (1000, 1.0), (300_007, 1.0)):
    gen = torch.Generator(device="cuda").manual_seed(P)
    enc = (torch.rand(P, 32, device="cuda", generator=gen) - 0.5).half()
    dirs = torch.randn(P, 3, device="cuda", generator=gen)
    dy = torch.randn(P, 4, device="cuda", generator=gen) * scale
    g32, de32 = ops.field_bwd(enc, dirs, 3, 1, wt, False, dy, True, [True] * 14)
    g16, de16 = ops.field_bwd(enc, dirs, 3, 1, wt, True, dy, True, [True] * 14)
    print(f"P={P} scale={scale}")
    for key, a, b in zip(synth.EXPERT_KEYS, g16, g32):
        print(f"   {key:32s} rel={rel(a, b):.4f}  |ref|={float(b.norm()):.3e} |got|={float(a.norm()):.3e}")
    print(f"   d_enc rel={rel(de16, de32):.4f}")
    # channel-wise: only rgb grads, only sigma grads
    for name, m in (("rgb-only", torch.tensor([1, 1, 1, 0.])), ("sigma-only", torch.tensor([0, 0, 0, 1.]))):
        dym = dy * m.cuda()
        g32, de32 = ops.field_bwd(enc, dirs, 3, 1, wt, False, dym, True, [True] * 14)
        g16, de16 = ops.field_bwd(enc, dirs, 3, 1, wt, True, dym, True, [True] * 14)
        print(f"   [{name}] " + " ".join(f"{rel(a, b):.3f}" for a, b in zip(g16, g32)) + f" denc={rel(de16, de32):.3f}")
