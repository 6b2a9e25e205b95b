import sys
from pathlib import Path
import torch
ROOT = Path("/root/repo")
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import synth
from helpers import cu
from adaptive_city_nerf_b200 import ops
P, S = 1 << 22, 64
sd = synth.make_expert_params(5, log2T=4)
wt = [cu(w) for w in synth.expert_weight_list(sd)]
enc = (torch.rand(P, 32, device="cuda") - 0.5)
rays = torch.randn(P // S, 8, device="cuda")
dy = torch.randn(P, 4, device="cuda") * 1e-3
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f = timeit(lambda: ops.field_fwd(enc, rays[:, 3:], 8, S, wt, False))
b = timeit(lambda: ops.field_bwd(enc, rays[:, 3:], 8, S, wt, False, dy, True, [True] * 14))
b2 = timeit(lambda: ops.field_bwd(enc, rays[:, 3:], 8, S, wt, False, dy, False, [True] * 14))
print(f"fp32 SIMT field at P={P}: fwd {f:.2f} ms, bwd {b:.2f} ms, bwd(no d_enc) {b2:.2f} ms -> per 2^24: {4*f:.1f} / {4*b:.1f} / {4*b2:.1f} ms")
