import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
dev = torch.device("cuda", 0)
net = util.build_network(dev); eng = net.engine()
xs = [util.synthetic_batch(32, 4096, start=32 * i).to(dev) for i in range(4)]
batches = [xs[i % 4] for i in range(40)]
with torch.no_grad():
    eng.forward_stream(batches[:8]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.forward_stream(batches)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/40:.3f} ms/batch, total {1e3*(t2-t0)/40:.3f} ms/batch")
