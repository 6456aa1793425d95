"""Shared helpers for the tests: seeded synthetic clouds and a deterministic, reference-free weight fill."""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

PATCHAUGNET_CFG = dict(  # configs/patch_aug_net.yaml:1-56 (model keys only)
    AGGREGATION="spvlad", AGGREGATION_TYPE=2, GROUP=8, NUM_POINTS=4096, FEATURE_OUTPUT_DIM=256,
    FEATURE_SIZE=[256, 256, 256], MAX_SAMPLES=[128, 1024, 4096], CLUSTER_SIZE=[4, 16, 64], OUTPUT_DIM=[256, 256, 256],
    USE_ORIGIN_PC_IN_FP=True, USE_SPA_ATT_AFTER_FP=True, GATING=False, SAMPLING=[1024, 128, 16], KNN=[20, 20, 20],
    KNN_DILATION=2)


PPTNET_CFG = dict(  # configs/pptnet_origin.yaml (model keys only)
    AGGREGATION="spvlad", GROUP=8, NUM_POINTS=4096, FEATURE_OUTPUT_DIM=256, FEATURE_SIZE=[256, 256, 256, 256],
    MAX_SAMPLES=[64, 256, 1024, 4096], CLUSTER_SIZE=[1, 4, 16, 64], OUTPUT_DIM=[256, 256, 256, 256], GATING=True,
    SAMPLING=[1024, 256, 64, 16], KNN=[20, 20, 20, 20])


def synthetic_cloud(i, n=4096):
    """Cloud i of SURVEY.md section 8d: seed 1234+i, uniform in [-1,1]^3, centred, scaled into the unit ball
    (utils/loading_pointclouds.py:51-63)."""
    g = torch.Generator().manual_seed(1234 + i)
    xyz = torch.rand(n, 3, generator=g) * 2 - 1
    xyz = xyz - xyz.mean(0, keepdim=True)
    xyz = xyz / xyz.norm(dim=1).max()
    return xyz.float()


def synthetic_batch(b, n=4096, start=0):
    return torch.stack([synthetic_cloud(start + i, n) for i in range(b)]).unsqueeze(1).contiguous()   # (B,1,N,3)


def tie_stress_cloud(i, n=4096):
    """5 % exact duplicates + zero padding rows (SURVEY.md section 8d 'tie stress set')."""
    xyz = synthetic_cloud(1000 + i, n).clone()
    g = torch.Generator().manual_seed(99 + i)
    ndup = n // 20
    src = torch.randint(0, n, (ndup,), generator=g)
    dst = torch.randint(0, n, (ndup,), generator=g)
    xyz[dst] = xyz[src]
    xyz[n - n // 50:] = 0.0
    return xyz


def base_place(p, n=8192):
    """Structured synthetic place p (after SURVEY.md section 8d): 2-8 random planes + 0-4 upright cylinders + 2-30 %
    uniform noise, in a place-specific anisotropic box (wide/flat like a street scene).  Places differ in gross shape
    and composition, which is what descriptors of a randomly initialised network can tell apart."""
    rng = np.random.default_rng(10_000 + p)
    npl, ncy = int(rng.integers(2, 9)), int(rng.integers(0, 5))
    noise_frac = rng.uniform(0.02, 0.3)
    scale = np.array([rng.uniform(0.25, 1.0), rng.uniform(0.05, 0.4), rng.uniform(0.25, 1.0)])
    per = int(n * (1 - noise_frac)) // (npl + ncy)
    pts = []
    for _ in range(npl):                                 # planes: random point + two in-plane axes + extents
        o = rng.uniform(-1, 1, 3)
        a, b = rng.normal(size=3), rng.normal(size=3)
        a /= np.linalg.norm(a); b -= a * (a @ b); b /= np.linalg.norm(b)
        uv = rng.uniform(-1, 1, (per, 2)) * rng.uniform(0.2, 0.9, 2)
        pts.append(o + uv[:, :1] * a + uv[:, 1:] * b)
    for _ in range(ncy):                                 # cylinders along the up (y) axis
        o = rng.uniform(-1, 1, 3)
        r = rng.uniform(0.05, 0.3)
        th = rng.uniform(0, 2 * np.pi, per)
        h = rng.uniform(-0.8, 0.8, per)
        pts.append(o + np.stack([r * np.cos(th), h, r * np.sin(th)], 1))
    pts = np.concatenate(pts)
    noise = rng.uniform(-1.2, 1.2, (n - len(pts), 3))
    return (np.concatenate([pts, noise]) * scale).astype(np.float32)


def place_visit(p, v, npts=4096, rotate=False):
    """Visit v of place p: (optional yaw, utils/loading_pointclouds.py:102-128) -> jitter N(0,0.005) clipped at 0.05
    (:163-174) -> random npts-subset -> unit-ball normalisation (:51-63).  Returns (npts,3) float32 numpy."""
    rng = np.random.default_rng(1_000_000 * (v + 1) + p)
    base = base_place(p)
    ang = rng.uniform(0, 2 * np.pi) if rotate else 0.0
    c, s = np.cos(ang), np.sin(ang)
    rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)          # rotate_point_cloud: about the up axis
    pts = base @ rot
    pts = pts + np.clip(0.005 * rng.normal(size=pts.shape), -0.05, 0.05).astype(np.float32)
    pts = pts[rng.choice(len(pts), npts, replace=False)]
    pts = pts - pts.mean(0, keepdims=True)                                    # normalize_point_cloud
    return (pts / np.max(np.linalg.norm(pts, axis=1))).astype(np.float32)


def place_batch(places, v, npts=4096):
    return torch.from_numpy(np.stack([place_visit(p, v, npts) for p in places])).unsqueeze(1).contiguous()


def golden_batch(model="patchaugnet"):
    """The 8 clouds of tests/golden/{patchaugnet,pptnet}_ref_forward.npz: two uniform clouds, one tie-stress cloud
    (exact duplicates + zero rows), five structured places."""
    u, t, p = (0, 0, 0) if model == "patchaugnet" else (10, 1, 5)
    return torch.cat([synthetic_batch(2, 4096, u), tie_stress_cloud(t)[None, None], place_batch(range(p, p + 5), 0)], 0)


def calibration_batch():
    """The clouds the BatchNorm running statistics of the test weights are calibrated on (make_calibration.py):
    32 structured places (ids 900000+) and 8 uniform clouds (ids 500+), none of which is used by a parity test."""
    return torch.cat([place_batch(range(900_000, 900_032), 0), synthetic_batch(8, 4096, start=500)], 0)


_CALIB = {}


def calibrated_bn(name):
    """BatchNorm running_mean / running_var fixture written by tests/golden/make_calibration.py: the statistics of one
    train-mode pass of the REFERENCE's module (weights = fill_state_dict) over calibration_batch().  With them every
    BatchNorm sees activations of the scale it was 'trained' on, so descriptors of different clouds differ by O(1)
    instead of collapsing onto the BN shift (round-1 VERDICT, weak #1)."""
    if name not in _CALIB:
        path = os.path.join(GOLDEN, f"calibrated_bn_{name}.npz")
        _CALIB[name] = {k: torch.from_numpy(v) for k, v in np.load(path).items()} if os.path.exists(path) else {}
    return _CALIB[name]


def fill_state_dict(sd, seed=123, calibrated=None):
    out = fill_state_dict_raw(sd, seed)
    if calibrated:
        cal = calibrated_bn(calibrated)
        for k, v in cal.items():
            if k in out and tuple(out[k].shape) == tuple(v.shape):
                out[k] = v.to(out[k].dtype).clone()
    return out


def fill_state_dict_raw(sd, seed=123):
    """Deterministic values for every entry of a state_dict, independent of module construction order and of the
    reference code: each tensor is drawn from a generator seeded by crc32(key) ^ seed."""
    out = {}
    bn_prefixes = {k[: -len("running_mean")] for k in sd if k.endswith("running_mean")}
    for k, v in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) ^ seed) & 0x7FFFFFFF)
        shape = tuple(v.shape)
        prefix = k[: k.rfind(".") + 1]
        if k.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif prefix in bn_prefixes and k.endswith("weight"):
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif prefix in bn_prefixes and k.endswith("bias"):
            t = torch.randn(shape, generator=g) * 0.1
        elif v.dim() >= 2:
            if "cluster_weights" in k or "hidden" in k or "gating_weights" in k:
                fan = shape[-2]
                t = torch.randn(shape, generator=g) / fan ** 0.5
            else:
                fan = int(np.prod(shape[1:]))
                t = torch.randn(shape, generator=g) * (2.0 / fan) ** 0.5
        else:
            t = torch.randn(shape, generator=g) * 0.05
        if k.endswith(("q_conv.weight", "k_conv.weight")):
            # PPT-Net attention (pptnet.py:246-282): with He-scaled q/k the Gram energies have std ~500, the row softmax
            # is an arg-max and the layer amplifies fp32 rounding ~20x per level (the reference's own fp32 forward is then
            # only reproducible to 2e-2 at the coarsest level).  0.05 puts the energies at O(1), like a trained network's.
            t = t * 0.05
        out[k] = t.to(v.dtype)
    for k in list(out):            # tied projections (pptnet.py:254): q_conv.weight IS k_conv.weight
        if k.endswith("q_conv.weight") and k[:-len("q_conv.weight")] + "k_conv.weight" in out:
            out[k] = out[k[:-len("q_conv.weight")] + "k_conv.weight"].clone()
    return out


def build_network(device="cpu", seed=123, cfg=None):
    from patchaugnet_b200.patch_aug_net import Network
    net = Network(param=dict(cfg or PATCHAUGNET_CFG), use_a2a_recon=True, use_l2_norm=True)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed, calibrated="patchaugnet"))
    return net.to(device).eval()


def build_pptnet(device="cpu", seed=321, cfg=None):
    from patchaugnet_b200.pptnet import Network
    net = Network(param=dict(cfg or PPTNET_CFG), use_normalize=True)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed, calibrated="pptnet"))
    return net.to(device).eval()
