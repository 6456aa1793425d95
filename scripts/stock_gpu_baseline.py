#!/usr/bin/env python
"""Stock-GPU proxy baseline (BASELINE.md section 4 "Stock GPU"): the reference's op-by-op forward structure with the
reference's OWN CUDA kernels (oracle/_ref, compiled from /root/reference for sm_100) for every pointops call and
PyTorch/cuDNN for conv/BN/matmul — i.e. what `place_recognition/patch_aug_net` does on a B200 with its stock
extensions.  MEASUREMENT TOOL (it loads oracle/_ref): not part of the product.

The module code is this repo's mirror of the reference modules run with `use_fused = False` (same op sequence as
patch_aug_net.py:203-243, 331-363); `patchaugnet_b200.pointops_cuda.*` is monkey-patched to the reference launchers.
The reference launchers run on the legacy default stream, so the whole forward is timed with device synchronisation
on both sides, like the reference's own timing (datasets/scene_dataset.py:672-686).
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import ctypes as C
    import util
    from oracle import refgpu
    from patchaugnet_b200 import pointops_cuda as K

    if not refgpu.available():
        print(json.dumps({"stock_gpu_proxy": None, "why": "oracle/_ref not built"}))
        return
    lib = refgpu._k()
    P = lambda t: C.c_void_p(t.data_ptr())
    # the reference's launchers use the legacy default stream: PyTorch's current stream here is the default stream too
    K.furthestsampling_cuda = lambda b, n, m, xyz, temp, idx: lib.furthestsampling_cuda_launcher(b, n, m, P(xyz), P(temp), P(idx))
    K.gathering_forward_cuda = lambda b, c, n, m, p, idx, out: lib.gathering_forward_cuda_launcher(b, c, n, m, P(p), P(idx), P(out))
    scratch = torch.zeros(1024, device="cuda")                # the reference writes dist2 un-offset (first nsample floats)
    K.knnquery_cuda = lambda b, n, m, ns, xyz, new_xyz, idx, d2: lib.knnquery_cuda_launcher(
        b, n, m, ns, P(xyz), P(new_xyz), P(idx), P(scratch), C.c_void_p(0))
    K.grouping_forward_cuda = lambda b, c, n, m, ns, p, idx, out: lib.grouping_forward_cuda_launcher_fast(b, c, n, m, ns, P(p), P(idx), P(out))
    K.nearestneighbor_cuda = lambda b, n, m, u, k, d2, idx: lib.nearestneighbor_cuda_launcher_fast(b, n, m, P(u), P(k), P(d2), P(idx))
    K.interpolation_forward_cuda = lambda b, c, m, n, p, idx, w, out: lib.interpolation_forward_cuda_launcher_fast(
        b, c, m, n, P(p), P(idx), P(w), P(out))

    from patchaugnet_b200 import pointops

    # the reference asks the kernel for knn_dilation * nsample neighbours (pointops.py:553); the mirror asks for nsample
    def stock_neighbour_idx(self, xyz, new_xyz):
        cand = pointops.knnquery(self.knn_dilation * self.nsample, xyz, new_xyz)
        perm = torch.randperm(self.nsample)
        return cand[:, :, perm.to(cand.device)].contiguous()
    pointops.QueryAndGroup_Edge.neighbour_idx = stock_neighbour_idx

    dev = torch.device("cuda", 0)
    results = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False        # torch default
        net = util.build_network(dev)
        net.use_fused = False
        x = util.synthetic_batch(32, 4096, start=0).to(dev)
        with torch.no_grad():
            for _ in range(2):
                net(x)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                net(x)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
        med = sorted(ts)[len(ts) // 2]
        results["cudnn_tf32" if tf32 else "fp32"] = dict(ms_per_batch=med * 1e3, submaps_per_s=32 / med)
    print(json.dumps({"stock_gpu_proxy": results, "batch": 32, "points": 4096,
                      "what": "reference op sequence + reference CUDA kernels (oracle/_ref, sm_100) + PyTorch conv/BN/matmul"}))


if __name__ == "__main__":
    main()
