"""Calibrate the BatchNorm running statistics of the TEST weights (build container only: needs /root/reference).

    python tests/golden/make_calibration.py          # then re-run make_golden.py

Round-1 weights (tests/util.fill_state_dict_raw) drew running_mean / running_var independently of the activations, so
the last BatchNorm's shift dominated the descriptor: different clouds gave descriptors that differed by 5e-3 and
retrieval was at chance.  Here ONE train-mode pass of the reference's own nn.Module code (the same CPU shims as
make_golden.py) over util.calibration_batch() — BatchNorm momentum None, i.e. running stats := the batch statistics of
that pass — yields statistics every layer actually sees.  They are committed as
tests/golden/calibrated_bn_{patchaugnet,pptnet}.npz and overlaid by util.fill_state_dict(..., calibrated=name).
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (sets sys.path for oracle / util)
import util  # noqa: E402

REF = make_golden.REF


def calibrate(net, x, chunks=1):
    for m in net.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = None
            m.reset_running_stats()
    net.train()
    torch.manual_seed(11)
    with torch.no_grad():
        net(x)
    net.eval()
    return {k: v.numpy().copy() for k, v in net.state_dict().items() if k.endswith(("running_mean", "running_var"))}


def main():
    make_golden.install_shims()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "place_recognition", "patch_aug_net", "models"))
    from place_recognition.patch_aug_net.models.patch_aug_net import Network
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "patch_aug_net.yaml")))
    net = Network(param=cfg, use_a2a_recon=True, use_l2_norm=True)
    net.load_state_dict(util.fill_state_dict_raw(net.state_dict(), seed=123))
    x = util.calibration_batch()
    stats = calibrate(net, x)
    np.savez_compressed(os.path.join(HERE, "calibrated_bn_patchaugnet.npz"), **stats)
    print("patchaugnet:", len(stats), "tensors,", sum(v.size for v in stats.values()), "floats")

    for m in [k for k in sys.modules if k == "loupe"]:
        del sys.modules[m]
    sys.path.insert(0, os.path.join(REF, "place_recognition", "pptnet_origin", "models"))
    from place_recognition.pptnet_origin.models.pptnet import Network as PPTNet
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "pptnet_origin.yaml")))
    net = PPTNet(param=cfg, use_normalize=True)
    net.load_state_dict(util.fill_state_dict_raw(net.state_dict(), seed=321))
    stats = calibrate(net, x)
    np.savez_compressed(os.path.join(HERE, "calibrated_bn_pptnet.npz"), **stats)
    print("pptnet:", len(stats), "tensors,", sum(v.size for v in stats.values()), "floats")


if __name__ == "__main__":
    main()
