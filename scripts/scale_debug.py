#!/usr/bin/env python
"""Why does the sharded extraction of bench.py's cfg4 leg not speed up on 2 GPUs?  Times the pieces on every rank."""
import os, sys, time, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import retrieval
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
net = util.build_network(dev); eng = net.engine()
g = torch.Generator(device=dev).manual_seed(1)
clouds = (torch.rand(6016, 4096, 3, generator=g, device=dev) * 2 - 1) * 0.57
def timed(label, fn):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rank {rank} {label}: host {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms", flush=True)
batches = [clouds[i:i + 32] for i in range(0, 6016, 32)]
with torch.no_grad():
    eng.forward_stream(batches[:4])
    timed("forward_stream 188 batches", lambda: eng.forward_stream(batches))
    timed("forward_stream 188 batches again", lambda: eng.forward_stream(batches))
    timed("extract_descriptors(6016 device clouds x world)", lambda: retrieval.extract_descriptors(net, clouds.repeat(1, 1, 1) if world == 1 else torch.cat([clouds] * world), batch_size=32, device=dev))
    big = torch.cat([clouds] * world)
    timed("extract_descriptors again", lambda: retrieval.extract_descriptors(net, big, batch_size=32, device=dev))
if world > 1: dist.destroy_process_group()
