"""Micro-benchmarks of individual C-ABI kernels (CUDA events, median of 20 after 3 warm-ups).  GPU box only."""
import ctypes as C
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from patchaugnet_b200 import _lib as L

lib = L.lib()
dev = "cuda"


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


res = {}
B = 32
xyz = (torch.rand(B, 4096, 3, device=dev) * 2 - 1).contiguous()
idx = torch.empty(B, 1024, dtype=torch.int32, device=dev)
for thr in (128, 256, 512, 1024):
    lib.pab_tune_fps_threads(thr)
    res[f"fps_4096_1024_t{thr}"] = timeit(lambda: lib.pab_furthestsampling(B, 4096, 1024, L.ptr(xyz), L.ptr(None), L.ptr(idx), L.stream_ptr()))
lib.pab_tune_fps_threads(0)
for bb in (32, 148, 296):
    x2 = (torch.rand(bb, 4096, 3, device=dev) * 2 - 1).contiguous()
    i2 = torch.empty(bb, 1024, dtype=torch.int32, device=dev)
    res[f"fps_4096_1024_B{bb}"] = timeit(lambda: lib.pab_furthestsampling(bb, 4096, 1024, L.ptr(x2), L.ptr(None), L.ptr(i2), L.stream_ptr()))
new_xyz = xyz[:, :1024].contiguous()
for k in (20, 40):
    nbr = torch.empty(B, 1024, k, dtype=torch.int32, device=dev)
    res[f"knn_4096_1024_k{k}"] = timeit(lambda: lib.pab_knnquery(B, 4096, 1024, k, L.ptr(xyz), L.ptr(new_xyz), L.ptr(nbr), L.ptr(None), L.stream_ptr()))
i3 = torch.empty(B, 4096, 3, dtype=torch.int32, device=dev)
w3 = torch.empty(B, 4096, 3, device=dev)
res["three_nn_4096_1024"] = timeit(lambda: lib.pab_three_nn_weights(B, 4096, 1024, L.ptr(xyz), L.ptr(new_xyz), L.ptr(i3), L.ptr(w3), L.stream_ptr()))
print(json.dumps(res, indent=1))
