"""Golden vectors for the input preparation: the reference's own normalize_point_cloud (utils/loading_pointclouds.py:51-63)
on seeded raw clouds.  Run in the build container (needs /root/reference): python tests/golden/make_prepare_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from utils.loading_pointclouds import normalize_point_cloud  # noqa: E402

rng = np.random.default_rng(2024)
out = {}
for i, (n, dtype) in enumerate([(2048, np.float64), (1000, np.float64), (512, np.float32)]):
    raw = (rng.normal(size=(n, 3)) * np.array([40.0, 25.0, 3.0]) + np.array([3.1e5, 4.2e6, 30.0])).astype(dtype)
    offset = np.array([3.1e5, 4.2e6, 0.0])
    pc = raw - offset
    for zoom in (True, False):
        res, meta = normalize_point_cloud(pc.copy(), True, zoom)
        out[f"raw{i}"] = raw
        out[f"offset{i}"] = offset
        out[f"pc{i}_zoom{int(zoom)}"] = res
        out[f"scale{i}_zoom{int(zoom)}"] = np.float64(meta["scale"])
        out[f"trans{i}_zoom{int(zoom)}"] = np.asarray(meta["trans"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prepare_ref.npz"), **out)
print("wrote prepare_ref.npz", {k: v.shape for k, v in out.items() if k.startswith("pc")})
