import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import _lib as L
g = np.load(os.path.join(util.GOLDEN, "pptnet_ref_forward.npz"))
dev = "cuda"
net = util.build_pptnet(dev)
x = torch.cat([util.synthetic_batch(1, 4096, 10), util.tie_stress_cloud(1)[None, None]], 0).to(dev)
def run(tag):
    with torch.no_grad():
        desc, fp, cidx = net(x)
    errs = [float(np.abs(fp[i][:, :, :8, 0].cpu().numpy() - g[f"fp{i}_head"]).max()) for i in range(4)]
    scl = [float(np.abs(g[f"fp{i}_head"]).max()) for i in range(4)]
    print(tag, "desc err %.2e" % np.abs(desc.cpu().numpy() - g["desc"]).max(), "fp errs", ["%.2e" % e for e in errs], "scales", ["%.2f" % s for s in scl])
    return desc, fp
run("fused engine, TC on ")
L.lib().pab_tune_tensor_core(0)
net._engine = None
d0, f0 = run("fused engine, TC off")
L.lib().pab_tune_tensor_core(1)
net.use_fused = False
run("module path        ")
