"""Extraction of device-resident clouds through retrieval.extract_descriptor_sets: per-call timing with / without coalescing."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import util
from patchaugnet_b200 import retrieval
dev = torch.device("cuda", 0)
net = util.build_network(dev)
g = torch.Generator(device=dev).manual_seed(1)
clouds = (torch.rand(6000, 4096, 3, generator=g, device=dev) * 2 - 1) * 0.57
for lb in (32, 128, 32, 128):
    for n in (256, 2048, 6000, 6000):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            d = retrieval.extract_descriptor_sets(net, [clouds[:n - 500], clouds[n - 500:n]], batch_size=32, device=dev, launch_batch=lb)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"launch_batch {lb:4d} clouds {n:5d}: {dt * 1e3:8.1f} ms  {n / dt:8.0f} clouds/s")
