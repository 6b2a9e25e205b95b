"""INTEGRATION.md section 1, checked against the real reference when it is mounted (build container only): after the
module aliasing recipe the reference's own callers -- nerfs/losses.py, pipelines/offline_stage/meta_core.py -- import
cleanly and their `render_rays` / `MetaContainer` ARE this package's.  Skipped where /root/reference does not exist."""
import importlib
import sys
import types
from pathlib import Path

import pytest

REF = Path("/root/reference")
ALIASES = ("models.encodings", "models.trunc_exp", "models.metamodule", "models.inr.meta_ngp", "models.inr.meta_container",
           "nerfs.scene_box", "nerfs.ray_sampling", "nerfs.ray_rendering")


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")
def test_reference_callers_bind_to_this_package():
    saved_modules = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        for name in list(sys.modules):                       # a clean slate for the reference's top-level packages
            if name.split(".")[0] in ("models", "nerfs", "pipelines", "common", "data", "utils"):
                del sys.modules[name]
        na = types.ModuleType("nerfacc"); na.OccGridEstimator = object
        sys.modules["nerfacc"] = na
        v = types.ModuleType("viser"); vt = types.ModuleType("viser.transforms"); v.transforms = vt
        sys.modules["viser"], sys.modules["viser.transforms"] = v, vt
        sys.path.insert(0, str(REF))
        import adaptive_city_nerf_b200  # noqa: F401
        for name in ALIASES:                                  # the recipe of INTEGRATION.md
            sys.modules[name] = importlib.import_module(f"adaptive_city_nerf_b200.{name}")
        losses = importlib.import_module("nerfs.losses")     # the reference's file, importing `.ray_rendering`
        ours = importlib.import_module("adaptive_city_nerf_b200.nerfs.ray_rendering")
        assert losses.render_rays is ours.render_rays
        assert Path(losses.__file__).is_relative_to(REF)
        meta_core = importlib.import_module("pipelines.offline_stage.meta_core")
        assert Path(meta_core.__file__).is_relative_to(REF) and callable(meta_core.task_adapt)
        from models.inr.meta_container import MetaContainer
        from adaptive_city_nerf_b200.models.inr.meta_container import MetaContainer as Ours
        assert MetaContainer is Ours
        # the expert keeps the interface extract_module_params relies on (meta_core.py:196-205)
        import torch
        from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
        aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
        m = Ours(num_submodules=1, centroids=torch.zeros(1, 3), aabb=aabb, use_bg_nerf=False, expert_box_list=[SceneBox(aabb)],
                 hidden=64, sigma_depth=2, color_depth=2, color_hidden=64, dir_encoding="spherical",
                 hash_enc_conf=dict(levels=4, features_per_level=2, log2_hashmap_size=8, max_res=64, min_res=16, interpolation="Linear"),
                 occ_conf={"use_occ": False})
        fast = meta_core.extract_module_params(m.submodules[0], copy=True)
        assert len(fast) == 14 and all(t.requires_grad for t in fast.values())
    finally:
        sys.path[:] = saved_path
        for name in list(sys.modules):
            if name not in saved_modules:
                del sys.modules[name]
        sys.modules.update(saved_modules)
