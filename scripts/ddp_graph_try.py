#!/usr/bin/env python
"""Does the CUDA-graph training step work under DistributedDataParallel?  (PyTorch: DDP built in a side stream, >= 11 warm-up
iterations before capture.)"""
import os, sys, time, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import training
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
dist.init_process_group("nccl", device_id=dev)
anchors = 16
net = util.build_network(dev).train()
model = training.build_ddp(net, dev, for_graph=True)
opt = torch.optim.Adam(model.parameters(), lr=5e-4, capturable=True)
g = torch.Generator(device=dev).manual_seed(77 + rank)
feed = (torch.rand(anchors * 18, 1, 4096, 3, generator=g, device=dev) * 2 - 1) * 0.57
step = training.GraphedTrainStep(model, opt, n_anchors=anchors, warmup=11)
t0 = time.perf_counter(); step(feed); torch.cuda.synchronize(); print(f"rank {rank} captured in {time.perf_counter()-t0:.1f} s", flush=True)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): loss, _ = step(feed)
e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: graphed DDP step {e0.elapsed_time(e1)/5:.1f} ms, loss {float(loss):.4f}", flush=True)
# weights identical on all ranks?
w = torch.cat([p.detach().flatten()[:1000] for p in net.parameters()])
lo, hi = w.clone(), w.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
print(f"rank {rank}: weights in sync across ranks: {bool(torch.equal(lo, hi))}", flush=True)
step.release()          # a live captured graph with NCCL nodes makes destroy_process_group hang
del step, model, opt
torch.cuda.synchronize()
os._exit(0)             # belt and braces: skip the NCCL teardown altogether
