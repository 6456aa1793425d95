"""FPS 4096 -> 1024 at B = 128: threads per cloud x clouds per CTA (machine time = SMs held x kernel time)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import util
from patchaugnet_b200 import _lib as L, pointops
lib = L.lib()
B = int(os.environ.get("FPS_B", 128))
x = torch.cat([util.synthetic_batch(16, 4096, 0)] * (B // 16)).squeeze(1).cuda().contiguous()
ref = None
for threads in (256, 128):
    for cpc in (1, 2):
        lib.pab_tune_fps_threads(threads)
        lib.pab_tune_fps_clouds_per_cta(cpc)
        try:
            for _ in range(3):
                idx = pointops.furthestsampling(x, 1024)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                idx = pointops.furthestsampling(x, 1024)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if ref is None:
                ref = idx.clone()
            per_sm = lib.pab_fps_clouds_per_sm(4096)
            print(f"threads {threads} clouds/CTA asked {cpc}: {ms:.3f} ms, identical {torch.equal(idx, ref)}")
        except Exception as ex:
            print(f"threads {threads} cpc {cpc}: {type(ex).__name__} {ex}")
lib.pab_tune_fps_threads(0); lib.pab_tune_fps_clouds_per_cta(1)
