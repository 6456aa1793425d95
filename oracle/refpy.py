"""The reference's OWN Python (oracle/_ref/refpy, copied unchanged by `make -C oracle refpy`) run on the GPU box over a
chosen ``pointops_cuda`` backend.  TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py) — GPU box only.

  backend "dropin"  <repo>/dropin/pointops_cuda.py, i.e. this repo's kernels behind the reference's pybind API
                    = "Option A" of INTEGRATION.md: libs/pointops/functions/pointops.py:8 does `import pointops_cuda`
                    and everything above it (QueryAndGroup_Edge, PointNet2, Network) is the reference's code.
  backend "stock"   the reference's own kernels (oracle/_ref/libref_kernels.so, compiled from /root/reference for
                    sm_100) behind the same 17-function API (pointops_api.cpp:15-40) = the stock libs/* build, bound
                    with ctypes instead of pybind.  This is the denominator of the north star's ">= 10x the stock
                    CUDA-extension build" and the strongest parity checker: the reference forward itself, on a B200.

The reference launchers run on the legacy default stream, which is PyTorch's default stream, so
no extra synchronisation is needed as long as the caller stays on the default stream.
"""
import ctypes as C
import importlib
import os
import sys
import types
import warnings

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REFPY = os.path.join(_HERE, "_ref", "refpy")
ROOT = os.path.dirname(_HERE)


def available():
    return os.path.exists(os.path.join(REFPY, "libs", "pointops", "functions", "pointops.py"))


def stock_pointops_cuda():
    """Module object with the pybind names of pointops_api.cpp:15-40 over the reference's compiled launchers."""
    from . import refgpu
    lib = refgpu._k()
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    m = types.ModuleType("pointops_cuda")
    null = C.c_void_p(0)        # cudaStream_t 0 = legacy default stream (what getCurrentCUDAStream() returns here)
    m.furthestsampling_cuda = lambda b, n, mm, xyz, temp, idx: lib.furthestsampling_cuda_launcher(b, n, mm, P(xyz), P(temp), P(idx))
    m.gathering_forward_cuda = lambda b, c, n, mm, p, idx, out: lib.gathering_forward_cuda_launcher(b, c, n, mm, P(p), P(idx), P(out))
    m.gathering_backward_cuda = lambda b, c, n, mm, g, idx, gp: lib.gathering_backward_cuda_launcher(b, c, n, mm, P(g), P(idx), P(gp))
    m.knnquery_cuda = lambda b, n, mm, ns, xyz, new_xyz, idx, d2: lib.knnquery_cuda_launcher(b, n, mm, ns, P(xyz), P(new_xyz), P(idx), P(d2), null)
    m.ballquery_cuda = lambda b, n, mm, r, ns, new_xyz, xyz, idx: lib.ballquery_cuda_launcher_fast(b, n, mm, C.c_float(r), ns, P(new_xyz), P(xyz), P(idx), null)
    m.grouping_forward_cuda = lambda b, c, n, mm, ns, p, idx, out: lib.grouping_forward_cuda_launcher_fast(b, c, n, mm, ns, P(p), P(idx), P(out))
    m.grouping_backward_cuda = lambda b, c, n, mm, ns, g, idx, gp: lib.grouping_backward_cuda_launcher(b, c, n, mm, ns, P(g), P(idx), P(gp))
    m.grouping_int_forward_cuda = lambda b, c, n, mm, ns, p, idx, out: lib.grouping_int_forward_cuda_launcher_fast(b, c, n, mm, ns, P(p), P(idx), P(out))
    m.nearestneighbor_cuda = lambda b, n, mm, u, k, d2, idx: lib.nearestneighbor_cuda_launcher_fast(b, n, mm, P(u), P(k), P(d2), P(idx))
    m.interpolation_forward_cuda = lambda b, c, mm, n, p, idx, w, out: lib.interpolation_forward_cuda_launcher_fast(b, c, mm, n, P(p), P(idx), P(w), P(out))
    m.interpolation_backward_cuda = lambda b, c, n, mm, g, idx, w, gp: lib.interpolation_backward_cuda_launcher(b, c, n, mm, P(g), P(idx), P(w), P(gp))
    return m


def dropin_pointops_cuda():
    """The module a user gets from `import pointops_cuda` with <repo>/dropin first on sys.path."""
    spec = importlib.util.spec_from_file_location("pointops_cuda_dropin", os.path.join(ROOT, "dropin", "pointops_cuda.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_LOADED = {}


def reference_modules():
    """Import the reference's pointops.py / patch_aug_net.py / pptnet.py from refpy (once).  Returns a namespace with
    .pointops, .patch_aug_net, .pptnet, .cfg_patchaugnet, .cfg_pptnet."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    import yaml
    if not available():
        raise RuntimeError("oracle/_ref/refpy missing: run `make -C oracle refpy` in the build container")
    if "pointops_cuda" not in sys.modules:
        sys.modules["pointops_cuda"] = dropin_pointops_cuda()      # pointops.py:8 binds the name at import
    warnings.filterwarnings("ignore", message=".*torch.cuda.*DtypeTensor constructors.*")
    sys.path.insert(0, REFPY)
    try:
        # both model directories put a top-level `loupe` / `pointnet_autoencoder` on sys.path (patch_aug_net.py:9-10);
        # import PatchAugNet first, then let PPT-Net see its own `loupe`
        pa = importlib.import_module("place_recognition.patch_aug_net.models.patch_aug_net")
        po = importlib.import_module("libs.pointops.functions.pointops")
        pa_loupe = sys.modules.pop("loupe")
        sys.path.insert(0, os.path.join(REFPY, "place_recognition", "pptnet_origin", "models"))
        pp = importlib.import_module("place_recognition.pptnet_origin.models.pptnet")
        sys.modules["loupe_pptnet"] = sys.modules.pop("loupe")
        sys.modules["loupe"] = pa_loupe
    finally:
        pass
    _LOADED.update(pointops=po, patch_aug_net=pa, pptnet=pp,
                   cfg_patchaugnet=yaml.safe_load(open(os.path.join(REFPY, "configs", "patch_aug_net.yaml"))),
                   cfg_pptnet=yaml.safe_load(open(os.path.join(REFPY, "configs", "pptnet_origin.yaml"))))
    return types.SimpleNamespace(**_LOADED)


_BACKENDS = {}


def use_backend(name):
    """Point the reference's pointops.py at `name` in {"dropin", "stock"} (its functions look the module global
    `pointops_cuda` up at call time, so swapping the global swaps every kernel)."""
    ref = reference_modules()
    if name not in _BACKENDS:
        _BACKENDS[name] = dropin_pointops_cuda() if name == "dropin" else stock_pointops_cuda()
    ref.pointops.pointops_cuda = _BACKENDS[name]
    return ref


def reference_patchaugnet(state_dict, device, backend):
    """The reference's Network (patch_aug_net.py:22-107) with the shipped YAML, given weights, eval mode, on `device`."""
    ref = use_backend(backend)
    net = ref.patch_aug_net.Network(param=ref.cfg_patchaugnet, use_a2a_recon=True, use_l2_norm=True)
    net.load_state_dict(state_dict)
    return net.to(device).eval()


def reference_pptnet(state_dict, device, backend):
    ref = use_backend(backend)
    net = ref.pptnet.Network(param=ref.cfg_pptnet, use_normalize=True)
    net.load_state_dict(state_dict)
    return net.to(device).eval()
