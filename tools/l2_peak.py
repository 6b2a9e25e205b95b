"""Measured ceilings of the memory system for the hash-grid access shape: scattered
8- / 16-byte loads and red.global.add vectors at pseudo-random rows of a buffer the size of the T = 2^19 table (64 MiB,
L2-resident) and of a buffer that spills L2 (1 GiB).  Prints one JSON object; bench.py reads profiles/*_l2_peaks.json.
    python tools/l2_peak.py [out.json]"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from adaptive_city_nerf_b200 import ops

dev = torch.device("cuda")
sm = torch.cuda.get_device_properties(dev).multi_processor_count
grid = sm * 8 * 4                       # 8 resident CTAs of 256 threads per SM, 4 waves
out = {"device": torch.cuda.get_device_name(dev), "sm_count": sm, "grid": grid, "block": 256, "rows": []}
for mib in (64, 1024):
    buf = torch.zeros(mib << 18, dtype=torch.float32, device=dev)
    for mode, name in ((0, "gather"), (1, "red")):
        for nbytes in (8, 16):
            iters = 16
            for _ in range(2):
                n = ops.debug_l2_probe(mode, nbytes, buf, iters, grid)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = ops.debug_l2_probe(mode, nbytes, buf, iters, grid)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out["rows"].append({"buffer_mib": mib, "op": name, "bytes": nbytes, "accesses": n, "ms": round(best, 4),
                                "g_accesses_per_s": round(n / best / 1e6, 2), "gb_per_s": round(n * nbytes / best / 1e6, 1)})
            if mode == 1:
                buf.zero_()
    del buf
print(json.dumps(out))
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
