"""Shared helpers for the parity tests."""
import numpy as np
import torch

import synth

F32 = np.float32


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def npy(t):
    return t.detach().float().cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, F32).view(np.uint32)


def assert_bitexact(a, b, what=""):
    a, b = np.asarray(a, F32), np.asarray(b, F32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: {np.count_nonzero(~same)} / {same.size} values differ bitwise (max abs {np.nanmax(np.abs(a - b))})"


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


HASH_CONF = dict(levels=16, features_per_level=2, log2_hashmap_size=12, max_res=4096, min_res=16, interpolation="Linear")


def make_container(K, centroids, boxes, margin, use_bg, seed0, hash_conf=None, cluster_2d=True, device="cuda"):
    """Our MetaContainer with the same seeded weights tests/golden/make_golden.py gives the reference."""
    from adaptive_city_nerf_b200.models.inr import MetaContainer
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    hash_conf = dict(hash_conf or HASH_CONF)
    T = torch.from_numpy
    m = MetaContainer(
        num_submodules=K, centroids=T(np.asarray(centroids, F32)), aabb=T(synth.AABB_GLOBAL),
        boundary_margin=margin, cluster_2d=cluster_2d, use_bg_nerf=use_bg,
        expert_box_list=[SceneBox(aabb=T(np.asarray(b, F32))) for b in boxes],
        hidden=64, sigma_depth=2, color_depth=2, color_hidden=64, dir_encoding="spherical",
        use_sigmoid_rgb=True, hash_enc_conf=hash_conf, occ_conf={"use_occ": False})
    sd = m.state_dict()
    for k in range(K):
        p = synth.make_expert_params(seed0 + k, L=hash_conf["levels"], F=hash_conf["features_per_level"],
                                     log2T=hash_conf["log2_hashmap_size"])
        for key, val in p.items():
            sd[f"submodules.{k}.{key}"] = T(val)
    if use_bg:
        for key, val in synth.make_bg_params(seed0 + 100).items():
            sd[key] = T(val)
    m.load_state_dict(sd)
    for k in range(K):  # the expert boxes are plain attributes: move them with the module
        m.submodules[k].scene_box = m.submodules[k].scene_box.to(device)
    return m.to(device)
