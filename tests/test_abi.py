"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/acn_b200.h declares, the ctypes table agrees, and the host mirror keeps the reference's
module surface (names, state_dict keys, params= rules) and refuses to compute on the CPU."""
import ctypes
import re
import subprocess
import warnings
from collections import OrderedDict
from pathlib import Path

import numpy as np
import pytest
import torch

import synth

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built_lib():
    from adaptive_city_nerf_b200 import build
    return build.build()


def header_symbols():
    text = (ROOT / "include" / "acn_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("acn_"))
    assert exported == header_symbols()


def test_debug_library_is_separate_and_complete(built_lib):
    """Probes and timelines live in libacn_b200_debug.so (include/acn_b200_debug.h): the product library exports no
    acn_debug_* symbol, the debug library exports every prototype of both headers."""
    from adaptive_city_nerf_b200 import _lib, build
    nm = lambda lib: subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    prod = [l.split()[-1] for l in nm(built_lib).splitlines() if " T " in l]
    assert not [s for s in prod if "debug" in s]
    dbg = {l.split()[-1] for l in nm(build.LIB_DEBUG).splitlines() if " T " in l}
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "acn_b200_debug.h").read_text(), flags=re.S)
    want = set(re.findall(r"\b(acn_debug_[a-z0-9_]+)\s*\(", text))
    assert len(want) >= 6 and want <= dbg and set(header_symbols()) <= dbg
    d = _lib.debug_lib()
    assert d.acn_version() == 200


def test_comm_library_exports_its_header(built_lib):
    """libacn_b200_comm.so (NCCL-backed exchange entries for non-PyTorch hosts) exports exactly what
    include/acn_b200_comm.h declares and binds through ctypes; no collective is called here (no GPU)."""
    from adaptive_city_nerf_b200 import build, comm
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "acn_b200_comm.h").read_text(), flags=re.S)
    want = sorted(set(re.findall(r"\b(acn_[a-z0-9_]+)\s*\(", text)))
    out = subprocess.run(["nm", "-D", "--defined-only", str(build.LIB_COMM)], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("acn_"))
    assert exported == want and len(want) == 8
    l = comm.comm_lib()
    assert l.acn_comm_rank(None, None, None) != 0 and b"null communicator" in l.acn_comm_last_error()
    assert len(l.acn_alltoall_samples.argtypes) == 7


def test_ctypes_table_matches_header(built_lib):
    from adaptive_city_nerf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    l = _lib.lib()
    assert l.acn_version() == 200
    for name in _lib.SIGNATURES:
        assert getattr(l, name).argtypes == _lib.SIGNATURES[name]
    assert len(_lib.SIGNATURES["acn_hashgrid_fwd"]) == 16 and len(_lib.SIGNATURES["acn_field_bwd"]) == 19


def test_error_convention_without_gpu(built_lib):
    from adaptive_city_nerf_b200 import _lib
    l = _lib.lib()
    if torch.cuda.is_available():
        pytest.skip("CPU-box check")
    out = ctypes.c_void_p()
    rc = l.acn_create(0, ctypes.byref(out))
    assert rc == -4 and b"no CUDA device" in l.acn_last_error()
    assert l.acn_destroy(None) == -1 and b"null context" in l.acn_last_error()


def test_product_refuses_cpu_tensors(built_lib):
    from adaptive_city_nerf_b200.models.encodings import HashGridEncoder, SHEncoder
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays, volume_render
    enc = HashGridEncoder(levels=2, log2_hashmap_size=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.rand(5, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        SHEncoder(levels=4)(torch.rand(5, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        volume_render(torch.rand(2, 4, 4), torch.rand(2, 4))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    for f in (ROOT / "adaptive_city_nerf_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h"):
            txt = f.read_text()
            assert "oracle" not in txt.replace("reference oracle", "").replace("oracle uses", "") or f.name == "ops.py", f
    txt = (ROOT / "adaptive_city_nerf_b200" / "ops.py").read_text()
    assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt


REFERENCE_KEYS = (
    ["scene_aabb_vec", "centroids"]
    + [f"submodules.{k}.{n}" for k in range(2) for n in
       ["aabb_extent", "xyz_encoder.hash_table"] + synth.EXPERT_KEYS]
    + ["bg_mlp.0.weight", "bg_mlp.0.bias", "bg_mlp.2.weight", "bg_mlp.2.bias"])


def _container(K=2, margin=1.05):
    from adaptive_city_nerf_b200.models.inr import MetaContainer
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    aabb = torch.from_numpy(synth.AABB_GLOBAL)
    return MetaContainer(
        num_submodules=K, centroids=torch.from_numpy(synth.CENTROIDS_G22[:K]), aabb=aabb, boundary_margin=margin,
        expert_box_list=[SceneBox(aabb)] * K, hidden=64, sigma_depth=2, color_depth=2, color_hidden=64,
        dir_encoding="spherical", use_sigmoid_rgb=True,
        hash_enc_conf=dict(levels=16, features_per_level=2, log2_hashmap_size=8, max_res=4096, min_res=16,
                           interpolation="Linear"), occ_conf={"use_occ": False})


def test_state_dict_keys_match_reference():
    m = _container()
    assert list(m.state_dict().keys()) == REFERENCE_KEYS          # SURVEY 5 "Checkpoint / resume"
    ex = m.submodules[0]
    assert ex.xyz_encoder.hash_table.shape == (16 * 256, 2) and ex.xyz_encoder.hash_table.dtype == torch.float32
    assert float(ex.sigma_head.bias) == -1.0                        # meta_ngp.py:83-84
    assert ex.xyz_encoder.hash_table.abs().max() <= 1e-3            # encodings.py:267
    assert list(ex.xyz_encoder.level_resolutions) == [16, 23, 33, 48, 70, 101, 147, 212, 307, 445, 645, 933, 1351, 1955, 2830, 4095]


def test_meta_parameters_and_param_groups():
    m = _container()
    names = [n for n, _ in m.meta_named_parameters()]
    assert names == [f"submodules.{k}.{n}" for k in range(2) for n in synth.EXPERT_KEYS]   # never the hash table / bg
    assert [n for n, _ in m.submodules[1].meta_named_parameters()] == synth.EXPERT_KEYS
    g = m.get_param_groups()
    assert set(g) == {"encoding", "sigma", "color", "background"}
    assert len(g["encoding"]["params"]) == 2 and len(g["sigma"]["params"]) == 16
    assert len(g["color"]["params"]) == 12 and len(g["background"]["params"]) == 4
    n_all = sum(p.numel() for p in m.parameters())
    assert sum(p.numel() for grp in g.values() for p in grp["params"]) == n_all


def test_get_subdict_semantics():
    m = _container()
    ex = m.submodules[0]
    fast = OrderedDict((n, p.detach().clone()) for n, p in m.meta_named_parameters())
    sub = m.get_subdict(fast, "submodules.1")
    assert list(sub) == synth.EXPERT_KEYS and sub["sigma_head.bias"] is fast["submodules.1.sigma_head.bias"]
    assert list(ex.get_subdict(sub, "sigma_trunk")) == ["0.linear.weight", "0.linear.bias", "1.linear.weight", "1.linear.bias"]
    assert m.get_subdict(None, "x") is None
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert ex.get_subdict(sub, "nope") is None
        assert any("has no parameter for submodule" in str(x.message) for x in w)
    # partial dicts fall back to own parameters tensor by tensor
    part = OrderedDict([("sigma_head.bias", torch.zeros(1))])
    ws = ex.fused_weights(part)
    assert ws[5] is part["sigma_head.bias"] and ws[0] is ex.sigma_trunk[0].linear.weight


def test_unsupported_configurations_raise():
    from adaptive_city_nerf_b200.models.inr import MetaNGP
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    box = SceneBox(torch.from_numpy(synth.AABB_GLOBAL))
    with pytest.raises(NotImplementedError):
        MetaNGP(occ_conf={"use_occ": True}, scene_box=box)
    with pytest.raises(NotImplementedError):
        MetaNGP(occ_conf={}, scene_box=box, dir_encoding="frequency")
    ex = MetaNGP(occ_conf={}, scene_box=box, hidden=128)
    with pytest.raises(NotImplementedError, match="fused field kernels"):
        ex._check_fused()


def test_linspace_table_is_reference_cpu_linspace(golden):
    from adaptive_city_nerf_b200 import ops
    g = golden("stage1")
    for S in (2, 3, 16, 17, 64, 65, 96, 255, 256):
        t = torch.linspace(0.0, 1.0, S)
        assert (t.numpy().view(np.uint32) == g[f"linspace_{S}"].view(np.uint32)).all()


def test_train_step_host_logic_without_gpu(built_lib):
    """Loss epilogue / optimizer tail: argument checks and the reference's get_optimizer grouping need no device; the
    struct the optimizer hands to the C ABI has the header's layout; compute refuses CPU tensors."""
    import ctypes as C
    import types
    from adaptive_city_nerf_b200 import _lib, ops
    from adaptive_city_nerf_b200.optim import FusedAdam, get_optimizer
    assert C.sizeof(_lib.AdamTensor) == 72                                 # 4 pointers + int64 + 2 doubles + step, bias pointers
    hdr = (ROOT / "include" / "acn_b200.h").read_text()
    assert f"#define ACN_ADAM_MAX_TENSORS {_lib.ADAM_MAX_TENSORS}" in hdr and f"#define ACN_LOSS_PARTIALS {_lib.LOSS_PARTIALS}" in hdr
    with pytest.raises(ValueError):
        ops.color_mse(torch.zeros(2, 3), torch.zeros(2, 3), "hsv")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.color_mse(torch.zeros(2, 3), torch.zeros(2, 3), "linear")

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.t, self.s, self.c, self.b = (torch.nn.Parameter(torch.zeros(n)) for n in (8, 4, 3, 2))

        def get_param_groups(self):
            return {"background": {"params": [self.b]}, "color": {"params": [self.c]}, "sigma": {"params": [self.s]},
                    "encoding": {"params": [self.t]}}

    P = types.SimpleNamespace(lr=5e-4, encoding_lr=1e-2, sigma_lr=None, color_lr=2e-3, optimizer="adamw", weight_decay=0.01)
    opt = get_optimizer(P, Holder())
    assert [(g["name"], g["lr"], g["weight_decay"]) for g in opt.param_groups] == [
        ("encoding", 1e-2, 0.01), ("sigma", 5e-4, 0.01), ("color", 2e-3, 0.01), ("background", 5e-4, 0.01)]
    assert opt.adamw and isinstance(opt, FusedAdam) and FusedAdam._step_supports_amp_scaling
    P.optimizer = "sgd"
    with pytest.raises(ValueError):
        get_optimizer(P, Holder())
    with pytest.raises(ValueError):
        FusedAdam([torch.nn.Parameter(torch.zeros(1))], betas=(1.0, 0.999))
    opt.param_groups[0]["params"][0].grad = torch.zeros(8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    opt2 = FusedAdam([torch.nn.Parameter(torch.zeros(1))])
    opt2.step()                                                             # nothing has a gradient: a no-op, as in torch


def test_task_grid_host_tensors_match_reference(golden):
    """data/task_binning.TaskGrid builds the region box, cell bounds and tolerances with torch ops (no kernel): they must
    equal what the reference's TaskDataset built (tests/golden/taskgrid.npz), bit for bit, on CPU tensors too."""
    from adaptive_city_nerf_b200.data import TaskGrid
    g = golden("taskgrid")
    rays = torch.from_numpy(synth.task_rays())
    for tag, cells, region in (("auto", (1, 6, 6), None), ("box", (2, 3, 4), tuple(map(tuple, synth.EXPERT_BOXES_G22[0].tolist())))):
        grid = TaskGrid(rays, cells, region)
        assert grid.num_cells == int(np.prod(cells))
        assert (grid.aabb.numpy().view(np.uint32) == g[f"{tag}_aabb"].view(np.uint32)).all()
        assert (grid.cell_bounds.numpy().view(np.uint32) == g[f"{tag}_cell_bounds"].view(np.uint32)).all()
        assert grid.cell3.shape == (3,) and grid.tol.shape == (grid.num_cells,) and bool((grid.tol >= 1e-9).all())


def test_integration_md_ctypes_example_matches_the_header():
    """The argument counts of the calls shown in INTEGRATION.md section 2 are the header's."""
    from adaptive_city_nerf_b200 import _lib
    text = (ROOT / "INTEGRATION.md").read_text()
    for name in ("acn_hashgrid_fwd_rays", "acn_field_fwd", "acn_render_expert_fwd", "acn_composite_fwd"):
        m = re.search(r"lib\." + name + r"\((.*?)\)\s*(?:#|\n[a-z#`])", text, flags=re.S)
        assert m, name
        args = re.sub(r"\([^()]*\)", "", m.group(1))              # drop nested calls such as C.c_int64(N)
        assert len([a for a in args.split(",") if a.strip()]) == len(_lib.SIGNATURES[name]), name
