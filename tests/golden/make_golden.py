"""Generate tests/golden/*.npz|json from the REFERENCE's own Python code (build container only: needs /root/reference).

    python tests/golden/make_golden.py

What runs here is the reference's unmodified nn.Module code (place_recognition/patch_aug_net/models/*.py,
libs/pointops/functions/pointops.py, utils/model_util/pt_util.py) on the CPU.  Two shims make that possible without
a GPU, and they are the only non-reference pieces on the path:
  * a module named ``pointops_cuda`` whose 17 functions are backed by the C oracle (oracle/pointops_oracle.c) — the
    oracle itself is pinned against the reference's compiled CUDA kernels on the GPU box (tests/golden/refgpu_*.npz,
    tests/test_refgpu.py);
  * ``torch.cuda.{Int,Float,Long}Tensor`` aliased to their CPU types (the reference allocates outputs with them).
Outputs (small): the state_dict manifest (key -> shape), and for seeded inputs + deterministically filled weights
(tests/util.py fill_state_dict) the reference descriptors, centre indices and feature checksums.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ops  # noqa: E402
import util  # noqa: E402


def install_shims():
    m = types.ModuleType("pointops_cuda")

    def wr(t, a):
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(t.dtype))

    m.furthestsampling_cuda = lambda b, n, mm, xyz, temp, idx: wr(idx, ops.furthestsampling(xyz.numpy(), mm, temp.numpy()))
    m.gathering_forward_cuda = lambda b, c, n, mm, p, idx, out: wr(out, ops.gathering(p.numpy(), idx.numpy()))
    m.knnquery_cuda = lambda b, n, mm, ns, xyz, new_xyz, idx, d2: wr(idx, ops.knnquery(ns, xyz.numpy(), new_xyz.numpy()))
    m.grouping_forward_cuda = lambda b, c, n, mm, ns, p, idx, out: wr(out, ops.grouping(p.numpy(), idx.numpy()))
    m.ballquery_cuda = lambda b, n, mm, r, ns, new_xyz, xyz, idx: wr(idx, ops.ballquery(r, ns, xyz.numpy(), new_xyz.numpy()))

    def nn3(b, n, mm, unknown, known, dist2, idx):
        d, i = ops.nearestneighbor(unknown.numpy(), known.numpy())
        wr(dist2, d); wr(idx, i)
    m.nearestneighbor_cuda = nn3
    m.interpolation_forward_cuda = lambda b, c, mm, n, p, idx, w, out: wr(out, ops.interpolation(p.numpy(), idx.numpy(), w.numpy()))
    for name in ["grouping_backward_cuda", "grouping_int_forward_cuda", "gathering_backward_cuda", "interpolation_backward_cuda",
                 "labelstat_idx_cuda", "labelstat_ballrange_cuda", "labelstat_and_ballquery_cuda", "featuredistribute_cuda",
                 "featuregather_forward_cuda", "featuregather_backward_cuda"]:
        setattr(m, name, None)
    sys.modules["pointops_cuda"] = m
    torch.cuda.IntTensor = torch.IntTensor
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.LongTensor = torch.LongTensor


def main():
    install_shims()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "place_recognition", "patch_aug_net", "models"))
    from place_recognition.patch_aug_net.models.patch_aug_net import Network  # the reference module

    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "patch_aug_net.yaml")))
    torch.manual_seed(123)
    net = Network(param=cfg, use_a2a_recon=True, use_l2_norm=True)
    manifest = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump({"n_params": sum(p.numel() for p in net.parameters()), "state_dict": manifest},
              open(os.path.join(HERE, "patchaugnet_state_dict.json"), "w"), indent=0)

    net.load_state_dict(util.fill_state_dict(net.state_dict(), seed=123))
    net.eval()
    B = 2
    x = torch.cat([util.synthetic_batch(1, 4096, 0), util.tie_stress_cloud(0)[None, None]], 0)   # one regular, one tie-stress
    torch.manual_seed(7)
    perms = []
    g = torch.get_rng_state()
    for _ in range(3):
        perms.append(torch.randperm(20).numpy())
    torch.set_rng_state(g)
    with torch.no_grad():
        desc, fp_features, center_idx = net(x)
    out = dict(desc=desc.numpy(), perms=np.stack(perms))
    for i, c in enumerate(center_idx):
        out[f"center_idx{i}"] = c.numpy().astype(np.int32)
    for i, f in enumerate(fp_features):
        f = f.numpy()
        out[f"fp{i}_sum"] = f.sum(axis=(2, 3)).astype(np.float64)          # (B,256) per-channel sums
        out[f"fp{i}_head"] = f[:, :, :8, 0].copy()                         # first 8 points, all channels
    np.savez_compressed(os.path.join(HERE, "patchaugnet_ref_forward.npz"), **out)
    print("desc[0,:6] =", desc[0, :6].numpy(), " |desc| =", desc.norm(dim=1).numpy())
    print("wrote", os.path.join(HERE, "patchaugnet_ref_forward.npz"))
    make_pptnet()


def make_pptnet():
    """Same procedure for the reference's PPT-Net (place_recognition/pptnet_origin/models/pptnet.py)."""
    for m in [k for k in sys.modules if k == "loupe"]:
        del sys.modules[m]                       # both model dirs ship a top-level module named `loupe`
    sys.path.insert(0, os.path.join(REF, "place_recognition", "pptnet_origin", "models"))
    from place_recognition.pptnet_origin.models.pptnet import Network as PPTNet
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "pptnet_origin.yaml")))
    torch.manual_seed(123)
    net = PPTNet(param=cfg, use_normalize=True)
    manifest = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump({"n_params": sum(p.numel() for p in net.parameters()), "state_dict": manifest},
              open(os.path.join(HERE, "pptnet_state_dict.json"), "w"), indent=0)
    net.load_state_dict(util.fill_state_dict(net.state_dict(), seed=321))
    net.eval()
    x = torch.cat([util.synthetic_batch(1, 4096, 10), util.tie_stress_cloud(1)[None, None]], 0)
    with torch.no_grad():
        desc, fp_features, center_idx = net(x)
    out = dict(desc=desc.numpy())
    for i, c in enumerate(center_idx):
        out[f"center_idx{i}"] = c.numpy().astype(np.int32)
    for i, f in enumerate(fp_features):
        f = f.numpy()
        out[f"fp{i}_sum"] = f.sum(axis=(2, 3)).astype(np.float64)
        out[f"fp{i}_head"] = f[:, :, :8, 0].copy()
    np.savez_compressed(os.path.join(HERE, "pptnet_ref_forward.npz"), **out)
    print("pptnet desc[0,:6] =", desc[0, :6].numpy(), " |desc| =", desc.norm(dim=1).numpy())


if __name__ == "__main__":
    main()
