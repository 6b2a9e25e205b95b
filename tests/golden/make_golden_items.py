"""Input builders shared by make_golden.py (which runs the reference on them) and the tests (which run the CUDA path)."""
import numpy as np

import synth


def ram_rays_items():
    """Duck-typed metadata (the attributes data/ram_rays_dataset.py reads from ImageMetadata) for five small synthetic
    images: plain, with a keep-mask, with a near/far that invalidates rows, an empty mask, a CHW uint8 image."""
    import types as _t
    rng = np.random.default_rng(401)
    items = []
    for i, cam in enumerate(synth.nadir_rays(402, 5, H=18, W=24, f=20.0)):
        img = rng.integers(0, 256, (18, 24, 3)).astype(np.uint8)
        mask = None
        if i == 1:
            mask = rng.uniform(0, 1, (18, 24)) > 0.4
        if i == 3:
            mask = np.zeros((18, 24), bool)
        if i == 2:                                                # outside the box, wide angle: part of the pixels miss it
            cam["c2w"] = synth.camera_c2w(1.25, 0.2, 0.3)
            cam["fx"] = cam["fy"] = 10.0
        items.append(dict(H=18, W=24, intrinsics=(cam["fx"], cam["fy"], cam["cx"], cam["cy"]), c2w=cam["c2w"], image_index=10 + i,
                          image=img if i != 4 else np.ascontiguousarray(img.transpose(2, 0, 1)), mask=mask))
    return items
