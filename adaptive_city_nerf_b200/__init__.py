"""adaptive_city_nerf_b200 -- B200-native (sm_100a) per-ray volumetric rendering hot path of
psklavos1/adaptive-city-nerf behind the reference's own Python module surface:

    adaptive_city_nerf_b200.models.encodings        <-> models/encodings.py
    adaptive_city_nerf_b200.models.metamodule       <-> models/metamodule/metamodule.py
    adaptive_city_nerf_b200.models.inr.meta_ngp     <-> models/inr/meta_ngp.py
    adaptive_city_nerf_b200.models.inr.meta_container <-> models/inr/meta_container.py
    adaptive_city_nerf_b200.nerfs.{scene_box,ray_sampling,ray_rendering} <-> nerfs/*.py

All arithmetic on the path runs in hand-written CUDA kernels reached through the C ABI of
libacn_b200.so (include/acn_b200.h).  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "1.0.0"
