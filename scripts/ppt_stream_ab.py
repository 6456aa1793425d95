"""PPT-Net throughput mode (batch 64 x 4096, 16 batches): static vs dynamic tensor-core tiles, fp32 contract and bf16 mode."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import util
dev = torch.device("cuda", 0)
ppt = util.build_pptnet(dev)
REP = int(os.environ.get("PPT_B", 64)) // 16
xs = [torch.cat([util.synthetic_batch(16, 4096, start=16 * j)] * REP).to(dev) for j in range(2)]
seq = [xs[i & 1] for i in range(16 * 4 // REP)]
for mode in ("f32", "bf16"):
    ppt.compute_dtype = mode
    eng = ppt.engine()
    for dyn in (True, True):
        eng.stream_dynamic_tiles = dyn
        with torch.no_grad():
            eng.forward_stream(seq[:4])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.forward_stream(seq)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / len(seq)
        print(f"{mode} batch {16 * REP} dynamic_tiles={dyn}: {ms:.3f} ms / batch, {16 * REP / ms * 1e3:.0f} submaps/s")
