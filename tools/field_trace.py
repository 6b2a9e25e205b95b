"""Timeline of the fused forward MLP kernel: row 0 of warpgroup 0 in CTA 0 (run on the GPU box)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import synth
from helpers import cu
from adaptive_city_nerf_b200 import ops
ops.use_debug_library()          # the timeline build of the kernels lives in libacn_b200_debug.so

P, S = 1 << 21, 64
sd = synth.make_expert_params(5, log2T=4)
wt = [cu(w) for w in synth.expert_weight_list(sd)]
enc = (torch.rand(P, 32, device="cuda") - 0.5).half()
rays = torch.randn(P // S, 8, device="cuda")
bwd = "--bwd" in sys.argv
dy = torch.randn(P, 4, device="cuda") * 1e-7
run = (lambda: ops.field_bwd(enc, rays[:, 3:], 8, S, wt, True, dy, True, [True] * 14)) if bwd else (lambda: ops.field_fwd(enc, rays[:, 3:], 8, S, wt, True))
run()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
ops.debug_field_trace(buf)
run()
torch.cuda.synchronize()
ops.debug_field_trace(None)
b = buf.cpu().tolist()
ev = [(v >> 8, v & 0xff) for v in b if v]
names = {1: "tile start", 2: "enc in TMEM + layer 1 issued", 3: "t0 done-wait returned", 4: "t0 epilogue done (TMEM -> TMEM)",
         5: "group barrier passed", 6: "layer 2 issued", 7: "t1 done-wait returned", 8: "heads done-wait returned",
         9: "c0 done-wait returned", 10: "c1 done-wait returned", 11: "c2 done-wait returned", 12: "tile finished"}
if bwd:
    names = {1: "tile start (xe staged)", 2: "barrier", 3: "F0 issued", 4: "F0 done", 5: "F0 epilogue done", 6: "barrier", 7: "F1 issued",
             8: "forward done, drr staged", 9: "barrier", 10: "B0 issued (dgrad + wgradT)", 11: "B0 dgrad done", 12: "mask computed",
             13: "B0 wgrad done", 14: "stored", 15: "barrier", 16: "B1 issued", 17: "tile finished"}
t0 = ev[0][0]
prev = t0
for t, tag in ev[:140]:
    print(f"  t={t - t0:7d} (+{t - prev:5d})  {tag:2d} {names.get(tag, '')}")
    prev = t
