"""Field-MLP kernels alone on synthetic encodings (run on the GPU box): timing at the bench size, or a short run
for ncu (`--small`)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import synth
from helpers import cu
from adaptive_city_nerf_b200 import ops

small = "--small" in sys.argv
P = (1 << 21) if small else (1 << 24)
S = 64
sd = synth.make_expert_params(5, log2T=4)
wt = [cu(w) for w in synth.expert_weight_list(sd)]
gen = torch.Generator(device="cuda").manual_seed(0)
enc = (torch.rand(P, 32, device="cuda", generator=gen) - 0.5).half()
rays = torch.randn(P // S, 8, device="cuda", generator=gen)
dirs = rays[:, 3:]
dy = torch.randn(P, 4, device="cuda", generator=gen) * 1e-7


def timeit(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = 1 if small else 5
f = timeit(lambda: ops.field_fwd(enc, dirs, 8, S, wt, True), n)
b = timeit(lambda: ops.field_bwd(enc, dirs, 8, S, wt, True, dy, True, [True] * 14), n)
b2 = timeit(lambda: ops.field_bwd(enc, dirs, 8, S, wt, True, dy, False, [True] * 14), n)
print(f"P={P} fwd {f:.3f} ms ({26880 * P / f / 1e9:.1f} TFLOP/s)  bwd {b:.3f} ms ({2 * 26880 * P / b / 1e9:.1f} TFLOP/s)  bwd(no d_enc) {b2:.3f} ms")
