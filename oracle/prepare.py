"""TEST INFRASTRUCTURE (oracle) — numpy restatement of the input preparation in front of the descriptor path:
SceneDataSet.get_pc (datasets/scene_dataset.py:713-740: load, subtract global_offset, normalise) and normalize_point_cloud
(utils/loading_pointclouds.py:51-63).  Pinned by tests/golden/prepare_ref.npz, produced by the reference's own
normalize_point_cloud (tests/golden/make_prepare_golden.py)."""
import numpy as np


def load_pc_file(path, dtype=np.float64):
    """utils/loading_pointclouds.py:14-24 (input_dim == 3 branch)."""
    return np.fromfile(path, dtype=dtype).reshape([-1, 3])


def normalize_point_cloud(pc, zoom=True):
    """utils/loading_pointclouds.py:51-63 -> (pc, {'scale': m, 'trans': centroid})."""
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = 1.0
    if zoom:
        m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
        pc = pc / m
    return pc, {"scale": m, "trans": centroid}


def get_pc(raw, global_offset, normalize=False, zoom=True):
    """scene_dataset.py:720-730 on an already loaded (N,3) array -> (pc float64, norm_meta)."""
    pc = raw - global_offset
    meta = {"scale": 1.0, "trans": np.zeros([1, 3])}
    if normalize:
        pc, meta = normalize_point_cloud(pc, zoom)
    return pc, meta
