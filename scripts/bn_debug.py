import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import pt_util, training
DEV = "cuda"
cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024])
feed = util.place_batch(range(700, 718), 0, 1024).to(DEV)
net = util.build_network(DEV, cfg=cfg).train()
orig = pt_util._BnReluTrain.apply
def checked(x, w, b, rm, rv, mom, eps):
    rm2, rv2 = rm.clone(), rv.clone()
    y = orig(x, w, b, rm, rv, mom, eps)
    with torch.no_grad():
        ref = torch.relu(torch.nn.functional.batch_norm(x.detach(), rm2, rv2, w.detach(), b.detach(), True, mom, eps))
        err = (ref - y).abs().max().item()
        print(tuple(x.shape), x.stride(), x.is_contiguous(), "fwd err %.2e" % err, "rm err %.2e" % (rm - rm2).abs().max().item(),
              "rv err %.2e" % (rv - rv2).abs().max().item(), "x absmax %.2e" % x.abs().max().item(), "var min %.2e" % x.detach().transpose(0,1).reshape(x.shape[1], -1).var(1).min().item())
    return y
pt_util._BnReluTrain.apply = staticmethod(checked)
torch.manual_seed(42)
x = feed.clone().requires_grad_(True)
desc, recon = net(x, training.make_nn_dict(1), return_feat=False)
