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

    net.load_state_dict(util.fill_state_dict(net.state_dict(), seed=123, calibrated="patchaugnet"))
    net.eval()
    x = util.golden_batch("patchaugnet")        # 8 clouds: 2 uniform, 1 tie-stress, 5 structured places
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
    d = desc.numpy()
    print("min over cloud pairs of max|desc_i - desc_j| =", min(np.abs(d[i] - d[j]).max() for i in range(8) for j in range(i)))
    print("wrote", os.path.join(HERE, "patchaugnet_ref_forward.npz"))
    make_decoder(net)
    make_pptnet()
    make_pointnetvlad()


def make_decoder(net):
    """PointNetDecoder (pointnet_autoencoder.py:85-111) of the reference Network built above: eval and train-mode
    values on seeded unit-norm patch features."""
    g = torch.Generator().manual_seed(77)
    f = torch.nn.functional.normalize(torch.randn(96, 256, generator=g))
    dec = net.decoder
    with torch.no_grad():
        dec.eval()
        out_eval = dec(f).numpy()
        dec.train()
        saved = {k: v.clone() for k, v in dec.state_dict().items()}
        out_train = dec(f).numpy()
        dec.load_state_dict(saved)
        dec.eval()
    np.savez_compressed(os.path.join(HERE, "decoder_ref.npz"), feats=f.numpy(), out_eval=out_eval, out_train=out_train)
    print("decoder:", out_eval.shape, float(np.abs(out_eval).mean()))


def make_pointnetvlad():
    """BASELINE.json configs[0]: the reference's PointNetVlad (pure PyTorch, runs on the CPU as is)."""
    sys.path.insert(0, os.path.join(REF, "place_recognition", "pointnet_vlad"))
    import importlib.util as iu
    spec = iu.spec_from_file_location("ref_pointnetvlad", os.path.join(REF, "place_recognition", "pointnet_vlad", "PointNetVlad.py"))
    mod = iu.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(123)
    net = mod.PointNetVlad(global_feat=True, feature_transform=True, max_pool=False, output_dim=256, num_points=4096)
    manifest = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump({"n_params": sum(p.numel() for p in net.parameters()), "state_dict": manifest},
              open(os.path.join(HERE, "pointnetvlad_state_dict.json"), "w"), indent=0)
    sd = util.fill_state_dict_raw(net.state_dict(), seed=55)
    net.load_state_dict(sd)
    # calibrate the BatchNorm statistics in place with one train-mode pass (same reasoning as make_calibration.py);
    # the statistics are a function of (seed, clouds) and are stored with the golden
    for m in net.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = None
            m.reset_running_stats()
    xc = torch.cat([util.place_batch(range(900_000, 900_012), 0), util.synthetic_batch(4, 4096, start=500)], 0)
    net.train()
    with torch.no_grad():
        net(xc)
    net.eval()
    stats = {k: v.numpy().copy() for k, v in net.state_dict().items() if k.endswith(("running_mean", "running_var"))}
    x = util.golden_batch("patchaugnet")[[0, 2, 3, 4]]
    with torch.no_grad():
        desc = net(x)
    np.savez_compressed(os.path.join(HERE, "pointnetvlad_ref_forward.npz"), desc=desc.numpy(),
                        **{"bn:" + k: v for k, v in stats.items()})
    d = desc.numpy()
    print("pointnetvlad desc", d.shape, "min pair diff", min(np.abs(d[i] - d[j]).max() for i in range(4) for j in range(i)))


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
    net.load_state_dict(util.fill_state_dict(net.state_dict(), seed=321, calibrated="pptnet"))
    net.eval()
    x = util.golden_batch("pptnet")
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
