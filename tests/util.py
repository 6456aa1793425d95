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


def fill_state_dict(sd, seed=123):
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
        out[k] = t.to(v.dtype)
    for k in list(out):            # tied projections (pptnet.py:254): q_conv.weight IS k_conv.weight
        if k.endswith("q_conv.weight") and k[:-len("q_conv.weight")] + "k_conv.weight" in out:
            out[k] = out[k[:-len("q_conv.weight")] + "k_conv.weight"].clone()
    return out


def build_network(device="cpu", seed=123, cfg=None):
    from patchaugnet_b200.patch_aug_net import Network
    net = Network(param=dict(cfg or PATCHAUGNET_CFG), use_a2a_recon=True, use_l2_norm=True)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed))
    return net.to(device).eval()


def build_pptnet(device="cpu", seed=321, cfg=None):
    from patchaugnet_b200.pptnet import Network
    net = Network(param=dict(cfg or PPTNET_CFG), use_normalize=True)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed))
    return net.to(device).eval()
