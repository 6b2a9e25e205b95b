"""How many SM cycles does a small tcgen05.mma cost?  Back-to-back MMAs per issuing thread, 1..4 issuing threads per SM
(run on the GPU box)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from adaptive_city_nerf_b200 import ops

print("mode  M   N  issuers | cycles per MMA at nmma = 1, 4, 16, 64 (per issuer, commit + wait included)")
for mode, name in ((0, "SS"), (1, "TS")):
    for M in (128, 64):
        for N in (64, 16, 8):
            if M == 128 and N == 8:
                continue
            for issuers in (1, 2, 4):
                row = []
                for nmma in (1, 4, 16, 64):
                    c = ops.debug_umma_rate(mode, M, N, nmma, 200, issuers)
                    row.append(max(c) / nmma)
                print(f"{name}   {M:3d} {N:3d}    {issuers}    | " + "  ".join(f"{v:7.1f}" for v in row))
