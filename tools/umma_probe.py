"""Hardware probe of tcgen05 operand layouts (run on the GPU box):  python tools/umma_probe.py
Each hypothesis stages known integer-valued matrices, runs the raw harness and compares the TMEM
dump with the expected product.  Output: gpurun_out/umma_probe.txt"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from adaptive_city_nerf_b200 import _lib  # noqa: E402

dev = torch.device("cuda")
lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    lines.append(s)


def idesc(M, N, a_bf16=False, b_bf16=False, a_mn=False, b_mn=False):
    return (1 << 4) | (int(a_bf16) << 7) | (int(b_bf16) << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(a, b, idsc, a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, ksteps, ncols):
    at = torch.from_numpy(a.view(np.uint16)).to(dev) if a.dtype == np.float16 else a
    bt = torch.from_numpy(b.view(np.uint16)).to(dev) if b.dtype == np.float16 else b
    out = torch.full((128, ncols), float("nan"), device=dev)
    l = _lib.debug_lib()
    _lib.debug_check(l.acn_debug_umma_raw(_lib.debug_ctx(dev), _lib.ptr(at), at.shape[0], at.shape[1], _lib.ptr(bt), bt.shape[0], bt.shape[1],
                                    idsc, a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, ksteps, ncols, _lib.ptr(out), _lib.stream(dev)))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def bf16_bits(x):
    return torch.from_numpy(x.astype(np.float32)).to(torch.bfloat16).view(torch.int16).to(dev)


rng = np.random.default_rng(0)


def ints(r, c):
    return rng.integers(-3, 4, (r, c)).astype(np.float16)


def report(name, got, exp, rows=slice(None)):
    ok = np.array_equal(got[rows], exp[rows])
    log(f"[{'OK' if ok else 'FAIL'}] {name}" + ("" if ok else f"  max|diff|={np.nanmax(np.abs(got[rows] - exp[rows]))} nan={np.isnan(got[rows]).sum()}"))
    return ok


K, N = 64, 64
A, W = ints(128, K), ints(N, K)
rgA, rgB = (K // 8) * 128, (K // 8) * 128
exp = A.astype(np.float32) @ W.astype(np.float32).T
Bs = ints(K, N)
rg = (N // 8) * 128
exp3 = A.astype(np.float32) @ Bs.astype(np.float32)
As = ints(K, 128)
rga = (128 // 8) * 128
exp4 = As.astype(np.float32).T @ W.astype(np.float32).T
G, X = ints(128, 64), ints(128, 64)
rg64 = (64 // 8) * 128
exp6 = G.astype(np.float32).T @ X.astype(np.float32)
X32 = ints(128, 32)
rg32 = (32 // 8) * 128
A64 = ints(64, K)
exp5 = A64.astype(np.float32) @ W.astype(np.float32).T


def lanes_of(got, expm):
    out = []
    for m in range(expm.shape[0]):
        hit = [l for l in range(128) if np.array_equal(got[l], expm[m])]
        out.append(hit[0] if len(hit) == 1 else (hit if hit else None))
    return out


HYP = {
    "H1": lambda: report("H1 K-major f16 x f16", run(A, W, idesc(128, N), 128, rgA, 256, 128, rgB, 256, K // 16, N), exp),
    "H2c": lambda: report("H2c A=bf16 B=bf16", run(bf16_bits(A), bf16_bits(W), idesc(128, N, a_bf16=True, b_bf16=True), 128, rgA, 256, 128, rgB, 256, K // 16, N), exp),
    "H3": lambda: report("H3  B MN-major (lbo=RG, sbo=128, step=2*RG)", run(A, Bs, idesc(128, N, b_mn=True), 128, rgA, 256, rg, 128, 2 * rg, K // 16, N), exp3),
    "H3p": lambda: report("H3' B MN-major (lbo=128, sbo=RG, step=2*RG)", run(A, Bs, idesc(128, N, b_mn=True), 128, rgA, 256, 128, rg, 2 * rg, K // 16, N), exp3),
    "H4": lambda: report("H4  A MN-major (lbo=RG, sbo=128, step=2*RG)", run(As, W, idesc(128, N, a_mn=True), rga, 128, 2 * rga, 128, rgB, 256, K // 16, N), exp4),
    "H4p": lambda: report("H4' A MN-major (lbo=128, sbo=RG, step=2*RG)", run(As, W, idesc(128, N, a_mn=True), 128, rga, 2 * rga, 128, rgB, 256, K // 16, N), exp4),
    "H6": lambda: report("H6  wgrad dW = G^T X both MN-major K=128 pts (rows 0..63)", run(G, X, idesc(128, 64, a_mn=True, b_mn=True), rg64, 128, 2 * rg64, rg64, 128, 2 * rg64, 8, 64)[:64], exp6),
    "H6p": lambda: report("H6' wgrad with lbo/sbo swapped", run(G, X, idesc(128, 64, a_mn=True, b_mn=True), 128, rg64, 2 * rg64, 128, rg64, 2 * rg64, 8, 64)[:64], exp6),
    "H6b": lambda: report("H6b wgrad N=32", run(G, X32, idesc(128, 32, a_mn=True, b_mn=True), rg64, 128, 2 * rg64, rg32, 128, 2 * rg32, 8, 32)[:64], G.astype(np.float32).T @ X32.astype(np.float32)),
    "H6c": lambda: report("H6c wgrad bf16 x bf16", run(bf16_bits(G), bf16_bits(X), idesc(128, 64, a_bf16=True, b_bf16=True, a_mn=True, b_mn=True), rg64, 128, 2 * rg64, rg64, 128, 2 * rg64, 8, 64)[:64], exp6),
    "H5": lambda: log("H5 M=64 K-major: row m -> TMEM lane:", lanes_of(run(A64, W, idesc(64, N), 128, rgA, 256, 128, rgB, 256, K // 16, N), exp5)),
    "H5b": lambda: log("H5b M=64 wgrad (MN-major): row m -> TMEM lane:", lanes_of(run(G, X, idesc(64, 64, a_mn=True, b_mn=True), rg64, 128, 2 * rg64, rg64, 128, 2 * rg64, 8, 64), exp6)),
    "H2a": lambda: report("H2a A=bf16 B=f16 (mixed)", run(bf16_bits(A), W, idesc(128, N, a_bf16=True), 128, rgA, 256, 128, rgB, 256, K // 16, N), exp),
}

if len(sys.argv) > 1:
    try:
        HYP[sys.argv[1]]()
    except Exception as e:  # noqa: BLE001
        log(f"[ERR] {sys.argv[1]}: {type(e).__name__}: {str(e).splitlines()[0]}")
    sys.exit(0)

import subprocess
allout = []
for name in HYP:
    r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
    txt = [l for l in r.stdout.splitlines() if l.startswith("[") or l.startswith("H5")]
    allout += txt or [f"[ERR] {name}: no output; stderr tail: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}"]
    print("\n".join(allout[-len(txt or [1]):]))
lines = allout
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "umma_probe.txt").write_text("\n".join(lines) + "\n")
