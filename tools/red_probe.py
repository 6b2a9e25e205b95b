"""RED rate of one B200 as a function of resident warps per SM, lane fill of the RED instructions and the arithmetic
between two REDs (acn_debug_red_probe): what a kernel with W warps per SM can push into a 64 MiB, L2-resident gradient
table.  Prints one JSON line per configuration.    python tools/red_probe.py [out.jsonl]"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from adaptive_city_nerf_b200 import ops

dev = torch.device("cuda")
sm = torch.cuda.get_device_properties(dev).multi_processor_count
l, chk, dctx = ops._dbg()
buf = torch.zeros(64 << 18, dtype=torch.float32, device=dev)
rows = []
for warps_per_sm in (64, 32, 16, 8, 4):
    for lanes in (32, 16, 8):
        for work in (0, 16, 64):
            block = min(1024, warps_per_sm * 32)
            grid = sm * (warps_per_sm * 32 // block)
            iters = max(1, 4096 // warps_per_sm // 8)

            def go():
                chk(l.acn_debug_red_probe(dctx(dev), ops.ptr(buf), buf.numel() * 4, iters, grid, block, lanes, work, ops.stream(dev)))
            go(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); go(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            n = grid * block // 32 * lanes * iters * 8
            ninstr = grid * block // 32 * iters * 8
            row = {"warps_per_sm": warps_per_sm, "active_lanes": lanes, "alu_ops_between_reds": work, "ms": round(best, 4),
                   "g_reds_per_s": round(n / best / 1e6, 1), "red_instr_per_sm_per_kcycle_at_1965MHz": round(ninstr / sm / (best * 1e-3 * 1.965e9) * 1e3, 2)}
            rows.append(row)
            print(json.dumps(row), flush=True)
            buf.zero_()
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text("\n".join(json.dumps(r) for r in rows) + "\n")
