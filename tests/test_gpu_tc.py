"""tcgen05 path: descriptor self-test and the fused fp16 field kernel vs the oracle's autocast
emulation and vs the fp32 kernels."""
import numpy as np
import pytest
import torch

import synth
from helpers import F32, cu, npy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(64, 64), (64, 32), (16, 64), (32, 16), (64, 16)])
def test_umma_tile_gemm(N, K):
    """D = A @ W^T on one tcgen05 tile (smem descriptors, instruction descriptor, TMEM layout)."""
    from adaptive_city_nerf_b200 import ops
    rng = np.random.default_rng(N * 100 + K)
    a = rng.integers(-4, 5, (128, K)).astype(np.float16)     # small integers: fp16 products and fp32 sums are exact
    w = rng.integers(-4, 5, (N, K)).astype(np.float16)
    d = ops.debug_umma_gemm(cu(a), cu(w))
    ref = a.astype(F32) @ w.astype(F32).T
    assert np.array_equal(npy(d), ref), f"max err {np.abs(npy(d) - ref).max()}"


@pytest.mark.parametrize("N,K", [(64, 64), (64, 32), (16, 64), (64, 16)])
def test_umma_tile_gemm_a_in_tmem(N, K):
    """Same product with the A operand written to tensor memory by tcgen05.st and read from there by the MMA."""
    from adaptive_city_nerf_b200 import ops
    rng = np.random.default_rng(7 + N * 100 + K)
    a = rng.integers(-4, 5, (128, K)).astype(np.float16)
    w = rng.integers(-4, 5, (N, K)).astype(np.float16)
    d = ops.debug_umma_gemm_ts(cu(a), cu(w))
    ref = a.astype(F32) @ w.astype(F32).T
    assert np.array_equal(npy(d), ref), f"max err {np.abs(npy(d) - ref).max()}"


def test_field_fp16_forward(golden, orc):
    from adaptive_city_nerf_b200 import ops
    g = golden("field")
    sd = synth.make_expert_params(31, log2T=12)
    ws = synth.expert_weight_list(sd)
    wt = [cu(w) for w in ws]
    enc16 = cu(g["enc"]).half()
    dirs = cu(g["dirs"])
    y = npy(ops.field_fwd(enc16, dirs, 3, 1, wt, half=True))
    ref_half = orc.field_fwd(npy(enc16), g["dirs"], ws, half=True)       # oracle's autocast emulation
    ref32 = g["y"]
    # north star: per-sample rgb within 1e-3 abs of the fp32 reference is the per-PIXEL bar after compositing;
    # per sample we allow fp16-rounding-level error: 4e-3 abs on rgb, 2e-2 rel on sigma
    assert np.abs(y[:, :3] - ref_half[:, :3]).max() < 2e-3
    assert np.abs(y[:, :3] - ref32[:, :3]).max() < 4e-3
    np.testing.assert_allclose(y[:, 3], ref32[:, 3], rtol=2e-2, atol=1e-4)
    # fp32 encodings are accepted too (converted in-kernel), ragged tile tails
    for P in (1, 127, 129, 1000):
        y2 = npy(ops.field_fwd(cu(g["enc"][:P]), dirs[:P].contiguous(), 3, 1, wt, half=True))
        assert np.abs(y2 - y[:P]).max() < 1e-6


def test_field_fp16_many_tiles_matches_fp32():
    """Persistent multi-CTA run: 300k points, every tile must agree with the fp32 kernel."""
    from adaptive_city_nerf_b200 import ops
    sd = synth.make_expert_params(5, log2T=4)
    wt = [cu(w) for w in synth.expert_weight_list(sd)]
    P = 300_007
    gen = torch.Generator(device="cuda").manual_seed(1)
    enc = (torch.rand(P, 32, device="cuda", generator=gen) - 0.5).half()
    dirs = torch.randn(P, 3, device="cuda", generator=gen)
    y16 = ops.field_fwd(enc, dirs, 3, 1, wt, half=True)
    y32 = ops.field_fwd(enc, dirs, 3, 1, wt, half=False)
    assert (y16[:, :3] - y32[:, :3]).abs().max() < 4e-3
    assert ((y16[:, 3] - y32[:, 3]).abs() / (y32[:, 3].abs() + 1e-3)).max() < 2e-2


def _rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _emulate_fp16_field(enc16, dirs, ws, dy):
    """Torch restatement of csrc/field_mma.cu's arithmetic (the reference's autocast path, models/metamodule/
    metamodule.py:150-155): fp16 operands, fp32 bias (a hi + lo fp16 pair), exact products with wide accumulation (fp64 here, fp32
    in TMEM), ReLU and one rounding to fp16; gradient tiles are fp16 carrying a power-of-two scale;
    the ReLU masks are those of THIS forward.  Returns (rgb_sigma, 14 grads, d_enc)."""
    from adaptive_city_nerf_b200 import ops
    D = torch.float64
    h = lambda t: t.half().to(D)                                   # round to fp16, continue exactly
    W = [w.to(D) for w in ws]
    wt0, bt0, wt1, bt1, wsg, bsg, wge, bge, wc0, bc0, wc1, bc1, wc2, bc2 = W
    G = wge.shape[0]
    x0 = enc16.to(D)
    hid = lambda x, w, b: h(torch.relu(x @ h(w).t() + b))         # fp32 bias (hi + lo fp16 pair) rides in the accumulator: ONE rounding, with ReLU
    h1 = hid(x0, wt0, bt0)
    h2 = hid(h1, wt1, bt1)
    sig_raw = (h2 @ h(wsg).t()).float().to(D) + bsg               # head accumulators stay fp32
    geo = (h2 @ h(wge).t()).float().to(D) + bge
    sh = ops.sh16(dirs).to(D)
    cin = torch.cat([h(geo), h(sh)], dim=1)
    c1 = hid(cin, wc0, bc0)
    c2 = hid(c1, wc1, bc1)
    y = torch.sigmoid((c2 @ h(wc2).t()).float().to(D) + bc2)
    e = torch.exp(sig_raw.clamp(-88.722839111, 88.722839111))
    out = torch.cat([y, e], dim=1).float()
    # ---- backward ----
    mx = float(dy.abs().max())
    scale = 1.0 if not (mx > 0) else 2.0 ** (10 - int(np.frexp(mx)[1]))
    dyd = dy.to(D)
    drr = h(dyd[:, :3] * scale * y * (1 - y))
    dsg = h(dyd[:, 3:] * scale * e)
    g = [None] * 14
    g[12], g[13] = drr.t() @ c2, drr.sum(0)
    gc2 = h(drr @ h(wc2)) * (c2 > 0)
    g[10], g[11] = gc2.t() @ c1, gc2.sum(0)
    gc1 = h(gc2 @ h(wc1)) * (c1 > 0)
    g[8], g[9] = gc1.t() @ cin, gc1.sum(0)
    dgeo = h((gc1 @ h(wc0))[:, :G])
    g[6], g[7] = dgeo.t() @ h2, dgeo.sum(0)
    g[4], g[5] = dsg.t() @ h2, dsg.sum(0)
    gh2 = h(dgeo @ h(wge) + dsg @ h(wsg)) * (h2 > 0)
    g[2], g[3] = gh2.t() @ h1, gh2.sum(0)
    gh1 = h(gh2 @ h(wt1)) * (h1 > 0)
    g[0], g[1] = gh1.t() @ x0, gh1.sum(0)
    d_enc = gh1 @ h(wt0)
    return out, [(t / scale).float().reshape(w.shape) for t, w in zip(g, ws)], (d_enc / scale).float()


@pytest.mark.parametrize("P,mag", [(1, 1.0), (128, 1e-7), (129, 1.0), (1000, 3e4), (300_007, 1e-7)])
def test_field_fp16_backward(P, mag):
    """tcgen05 backward (fp16 operands with a power-of-two loss scale, fp32 accumulate, TMEM-resident weight
    gradients) against (a) a torch restatement of the same arithmetic -- tight: relative L2 <= 3e-3 per tensor --
    and (b) the fp32 SIMT backward.  For (b) the gradients of a ReLU network evaluated in fp16 and in fp32 differ
    mostly through the ~0.1 % of units whose pre-activation changes sign under fp16 rounding; with independent
    random dL/dy that alone gives ~1-2 % relative L2 (measured with plain torch as well), hence the 4e-2 bar here
    while the coherent-gradient case (tests/test_gpu_render.py) keeps 2e-2."""
    from adaptive_city_nerf_b200 import ops
    sd = synth.make_expert_params(5, log2T=4)
    wt = [cu(w) for w in synth.expert_weight_list(sd)]
    gen = torch.Generator(device="cuda").manual_seed(P)
    enc = (torch.rand(P, 32, device="cuda", generator=gen) - 0.5).half()
    dirs = torch.randn(P, 3, device="cuda", generator=gen)
    # per-sample gradients of a mean loss are tiny (1e-7), GradScaler-scaled ones are huge (3e4): both must work
    dy = torch.randn(P, 4, device="cuda", generator=gen) * mag
    dy[::5] = 0.0
    g16, de16 = ops.field_bwd(enc, dirs, 3, 1, wt, True, dy, True, [True] * 14)
    y_em, g_em, de_em = _emulate_fp16_field(enc, dirs, wt, dy)
    y16 = ops.field_fwd(enc, dirs, 3, 1, wt, half=True)
    # same arithmetic up to the accumulation order: a hidden unit lands on the other side of an fp16 rounding
    # boundary now and then (one ulp), everything else agrees to MUFU precision
    assert (y16[:, :3] - y_em[:, :3]).abs().max() < 1e-3 and (y16[:, :3] - y_em[:, :3]).abs().mean() < 1e-5
    rel_s = (y16[:, 3] - y_em[:, 3]).abs() / (y_em[:, 3].abs() + 1e-6)
    assert rel_s.max() < 5e-3 and rel_s.mean() < 1e-5
    for key, a, b in zip(synth.EXPERT_KEYS, g16, g_em):
        assert torch.isfinite(a).all(), key
        assert _rel_l2(a, b) < 3e-3, ("vs fp16 restatement", key, _rel_l2(a, b))
    assert _rel_l2(de16, de_em) < 3e-3
    if P >= 1000:            # a statistical statement: with a few hundred points one flipped unit moves it by percents
        g32, de32 = ops.field_bwd(enc, dirs, 3, 1, wt, False, dy, True, [True] * 14)
        for key, a, b in zip(synth.EXPERT_KEYS, g16, g32):
            assert _rel_l2(a, b) < 4e-2, ("vs fp32", key, _rel_l2(a, b))
        assert _rel_l2(de16, de32) < 4e-2
    # skipping the input gradient (inner-loop steps) must not change the weight gradients
    g16b, none = ops.field_bwd(enc, dirs, 3, 1, wt, True, dy, False, [True] * 14)
    assert none is None
    for a, b in zip(g16b, g16):
        assert _rel_l2(a, b) < 1e-5
    # only some gradients requested
    need = [i % 2 == 0 for i in range(14)]
    g16c, _ = ops.field_bwd(enc, dirs, 3, 1, wt, True, dy, False, need)
    assert all((x is None) == (not n) for x, n in zip(g16c, need))
    for a, b, n in zip(g16c, g16, need):
        if n:
            assert _rel_l2(a, b) < 1e-5


def test_field_fp16_backward_zero_and_nonfinite_gradients():
    """All-zero dL/dy gives exactly zero gradients; the loss scale ignores NaN/inf entries (it falls back to 1)."""
    from adaptive_city_nerf_b200 import ops
    sd = synth.make_expert_params(5, log2T=4)
    wt = [cu(w) for w in synth.expert_weight_list(sd)]
    P = 777
    enc = (torch.rand(P, 32, device="cuda") - 0.5).half()
    dirs = torch.randn(P, 3, device="cuda")
    g, de = ops.field_bwd(enc, dirs, 3, 1, wt, True, torch.zeros(P, 4, device="cuda"), True, [True] * 14)
    assert all(float(t.abs().max()) == 0.0 for t in g) and float(de.abs().max()) == 0.0
    dy = torch.randn(P, 4, device="cuda")
    dy[3, 1] = float("inf")
    g, de = ops.field_bwd(enc, dirs, 3, 1, wt, True, dy, True, [True] * 14)
    assert not torch.isfinite(g[12]).all()          # the bad sample poisons what it touches, as in the reference


@pytest.mark.parametrize("E,G,mode", [(16, 7, "points"), (32, 1, "points"), (32, 15, "rays"), (16, 12, "rays")])
def test_field_fp16_other_widths_and_ray_mode(E, G, mode):
    """Encoding width 16 / geo width != 15 (runtime G: the colour-input columns are permuted in the kernels), and the
    ray-indexed direction mode (one direction per S consecutive points) against the per-point mode and the torch
    restatement."""
    from adaptive_city_nerf_b200 import ops
    sd = synth.make_expert_params(11, E=E, G=G, log2T=4)
    wt = [cu(w) for w in synth.expert_weight_list(sd)]
    S, N = 48, 211
    P = N * S
    gen = torch.Generator(device="cuda").manual_seed(E * 100 + G)
    enc = (torch.rand(P, E, device="cuda", generator=gen) - 0.5).half()
    rays = torch.randn(N, 8, device="cuda", generator=gen)
    dirs_pt = rays[:, 3:6].repeat_interleave(S, dim=0).contiguous()
    dy = torch.randn(P, 4, device="cuda", generator=gen) * 1e-3
    if mode == "rays":
        y = ops.field_fwd(enc, rays[:, 3:], 8, S, wt, half=True)
        g, de = ops.field_bwd(enc, rays[:, 3:], 8, S, wt, True, dy, True, [True] * 14)
        y_pt = ops.field_fwd(enc, dirs_pt, 3, 1, wt, half=True)
        assert torch.equal(y, y_pt)                       # same arithmetic, only the direction addressing differs
    else:
        y = ops.field_fwd(enc, dirs_pt, 3, 1, wt, half=True)
        g, de = ops.field_bwd(enc, dirs_pt, 3, 1, wt, True, dy, True, [True] * 14)
    y_em, g_em, de_em = _emulate_fp16_field(enc, dirs_pt, wt, dy)
    assert (y[:, :3] - y_em[:, :3]).abs().max() < 1e-3
    for key, a, b in zip(synth.EXPERT_KEYS, g, g_em):
        assert a.shape == b.shape and torch.isfinite(a).all(), key
        # 10 k points: one hidden unit on the other side of an fp16 rounding boundary (fp32 accumulation in TMEM vs fp64
        # here) moves a first-layer gradient by ~5e-3; the 300 k-point case of test_field_fp16_backward keeps 3e-3
        assert _rel_l2(a, b) < 8e-3, (key, _rel_l2(a, b))
    assert _rel_l2(de, de_em) < 8e-3
    y32 = ops.field_fwd(enc, dirs_pt, 3, 1, wt, half=False)
    assert (y[:, :3] - y32[:, :3]).abs().max() < 4e-3


@pytest.mark.parametrize("mode,L,log2T,interp", [("rays", 16, 14, "Linear"), ("points", 16, 12, "Smoothstep"), ("rays", 8, 10, "Linear")])
def test_fused_expert_backward_matches_two_kernel_path(mode, L, log2T, interp):
    """acn_render_expert_bwd (MLP backward + table scatter in one warp-specialised kernel, d_enc never in HBM) against acn_field_bwd ->
    d_enc -> acn_hashgrid_bwd[_rays]: the MLP arithmetic is the same kernel code, so the weight gradients agree to the
    order of the atomics and the table gradient to fp32 summation order (2e-5 relative L2; 1e-6 of the largest entry)."""
    from adaptive_city_nerf_b200 import ops, _lib
    E = 2 * L
    sd = synth.make_expert_params(17, E=E, log2T=4)
    wt = [cu(w) for w in synth.expert_weight_list(sd)]
    gen = torch.Generator(device="cuda").manual_seed(L + log2T)
    S, N = 48, 301
    P = N * S
    res = torch.tensor(np.floor(16 * np.exp(np.arange(L) * np.log(4096 / 16) / max(L - 1, 1))).astype(np.int32))
    spec = ops.GridSpec(L, 2, log2T, res, _lib.INTERP[interp])
    table = (torch.rand(L << log2T, 2, device="cuda", generator=gen) - 0.5) * 0.2
    lo, hi = synth.AABB_GLOBAL
    box6 = cu(np.concatenate([lo, hi - lo]).astype(F32))
    o = torch.rand(N, 3, device="cuda", generator=gen) * cu(hi - lo) * 0.2 + cu(lo) + cu(hi - lo) * 0.1
    d = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=gen), dim=-1)
    rays = torch.cat([o, d, torch.zeros(N, 1, device="cuda"), torch.ones(N, 1, device="cuda")], 1).contiguous()
    t = (torch.rand(N, S, device="cuda", generator=gen) * 0.5).sort(dim=1).values.contiguous()
    dy = torch.randn(P, 4, device="cuda", generator=gen) * 1e-6
    dy[::7] = 0.0
    if mode == "rays":
        enc = ops.hashgrid_fwd_rays(rays, t, table, spec, box6, torch.float16)
        pos, dirs, ds, dg = (rays, t), rays[:, 3:], 8, S
    else:
        x6 = ops.points(rays, t)
        enc = ops.hashgrid_fwd(x6, table, spec, box6, torch.float16)
        pos, dirs, ds, dg = (x6,), x6[:, 3:], 6, 1
    g2, d_enc = ops.field_bwd(enc, dirs, ds, dg, wt, True, dy, True, [True] * 14)
    dt2 = torch.zeros_like(table)
    if mode == "rays":
        ops.hashgrid_bwd_rays(rays, t, d_enc, spec, box6, dt2)
    else:
        ops.hashgrid_bwd(x6, d_enc, spec, box6, dt2)
    dt1 = torch.zeros_like(table)
    g1 = ops.render_expert_bwd(enc, pos, dirs, ds, dg, wt, dy, [True] * 14, spec, box6, dt1)
    for key, a, b in zip(synth.EXPERT_KEYS, g1, g2):
        assert _rel_l2(a, b) < 1e-5, (key, _rel_l2(a, b))
    assert float(dt2.abs().max()) > 0
    assert _rel_l2(dt1, dt2) < 2e-5, _rel_l2(dt1, dt2)
    assert float((dt1 - dt2).abs().max()) < 1e-6 * float(dt2.abs().max()) + 1e-30
    # the single-role fused kernel of the debug library (the shipped kernel's predecessor and A/B partner): same results
    dt5 = torch.zeros_like(table)
    g5 = ops.debug_render_expert_bwd_single(enc, pos, dirs, ds, dg, wt, dy, [True] * 14, spec, box6, dt5)
    for key, a, b in zip(synth.EXPERT_KEYS, g1, g5):
        assert _rel_l2(a, b) < 1e-5, (key, _rel_l2(a, b))
    assert _rel_l2(dt1, dt5) < 2e-5, _rel_l2(dt1, dt5)
    # a ragged tail (P not a multiple of the 128-point tile) and only some weight gradients requested
    Pq = P - 37 * S if mode == "rays" else P - 1777
    need = [i % 3 != 0 for i in range(14)]
    posq = (rays[:Pq // S], t[:Pq // S]) if mode == "rays" else (x6[:Pq],)
    dt3 = torch.zeros_like(table)
    g3 = ops.render_expert_bwd(enc[:Pq], posq, dirs, ds, dg, wt, dy[:Pq], need, spec, box6, dt3)
    g4, de4 = ops.field_bwd(enc[:Pq], dirs, ds, dg, wt, True, dy[:Pq], True, need)
    dt4 = torch.zeros_like(table)
    if mode == "rays":
        ops.hashgrid_bwd_rays(*posq, de4, spec, box6, dt4)
    else:
        ops.hashgrid_bwd(posq[0], de4, spec, box6, dt4)
    assert all((a is None) == (not n) for a, n in zip(g3, need))
    assert _rel_l2(dt3, dt4) < 2e-5
