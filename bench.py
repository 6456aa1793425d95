#!/usr/bin/env python
"""bench.py — 4096-pt submap descriptors/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32]

One "step" = one pass of the descriptor-extraction hot path (PatchAugNet eval forward, fp32) over one batch of
32 synthetic 4096-point submaps per GPU (BASELINE.json configs[1]).  Prints ONE JSON line on rank 0.

  value     device-resident throughput: inputs already in HBM, CUDA events around exactly K steps, max over ranks.
            Inputs rotate through a pool larger than the 126 MB L2 ("inputs_larger_than_l2").
  e2e       the same metric through the public nn.Module call (`net(x)`) with PINNED HOST input: the H2D copy of the
            batch and the D2H read of the (B,256) descriptors are inside the timed region.
  roofline  the dominant kernel of the step (chosen from per-stage CUDA events), timed live with events in the same
            timed region; algorithmic FLOPs/bytes per launch from patchaugnet_b200.engine.stage_work.
  cpu_baseline  the CPU oracle (oracle/, a port of the reference path) on a bounded sample, rank 0, N=1 only.

  stock_gpu     the reference's own Python + its own CUDA kernels compiled for sm_100 (oracle/_ref), same batch: the
                north star's ">= 10x the stock libs/* build" denominator, fp32 and cuDNN-TF32.
  pointnetvlad_cpu  BASELINE.json configs[0]: the reference's pure-PyTorch PointNetVLAD forward on the host cores.
  measured_peaks    fp32 / TF32 / bf16 matmul 8192^3 and copy bandwidth measured in this run (the fp32 configs' peaks).
  configs       sub-records for BASELINE.json configs[2] (PPT-Net batch 64) and configs[3] (10k-submap retrieval).

--impl reference: the reference path's CPU implementation (the oracle port: the reference ships no CPU code for
pointops, and its CUDA kernels are what this repo replaces) on the host cores, same metric / config / step shape
(32 clouds per step, scans OpenMP-parallel over all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "4096-pt submap descriptors/sec"
UNIT = "submaps/s"
NPTS = 4096


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")   # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def cpu_oracle_throughput(budget_s=15.0, clouds_per_call=32):
    """The oracle port of the path on the host cores: submaps/s over a bounded sample (about `budget_s` seconds)."""
    import util
    from oracle import model
    net = util.build_network("cpu")
    sd = net.state_dict()
    x = util.synthetic_batch(clouds_per_call, NPTS, start=0).numpy()
    model.patchaugnet_forward(sd, util.PATCHAUGNET_CFG, x[:1])           # warm-up (library load, thread pools)
    done, t0 = 0, time.perf_counter()
    while True:
        model.patchaugnet_forward(sd, util.PATCHAUGNET_CFG, x)
        done += clouds_per_call
        el = time.perf_counter() - t0
        if el >= budget_s or done >= 256:
            break
    return done / el, done, el


def run_reference(args):
    """--impl reference: oracle port on the host cores, same step shape as the GPU arm (args.batch clouds per step);
    rank 0 only.  Steps are capped so that the whole run stays within a few minutes on a slow host."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    per_step = args.batch
    import util
    from oracle import model
    net = util.build_network("cpu")
    sd = net.state_dict()
    x = util.synthetic_batch(per_step, NPTS, start=0).numpy()
    t0 = time.perf_counter()
    model.patchaugnet_forward(sd, util.PATCHAUGNET_CFG, x)          # warm-up 1 (library load, thread pools)
    t_first = time.perf_counter() - t0
    for _ in range(max(0, min(args.warmup, 2) - 1)):
        model.patchaugnet_forward(sd, util.PATCHAUGNET_CFG, x)
    steps = max(1, min(args.steps, int(180.0 / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        model.patchaugnet_forward(sd, util.PATCHAUGNET_CFG, x)
    el = time.perf_counter() - t0
    val = steps * per_step / el
    sample = (f"{steps} steps x {per_step} clouds x {NPTS} pts, oracle PatchAugNet eval forward, fp32 (C scans OpenMP over "
              f"{os.cpu_count()} threads, dense layers torch-CPU)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"PatchAugNet descriptor extraction, batch {per_step} x {NPTS}-pt synthetic clouds, fp32, eval "
                               "(BASELINE.json configs[1]) on the host CPU", "global_batch": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def measure_matmul_peaks(dev):
    """Same method as MEASURED_PEAKS.json (torch.matmul 8192^3, best of 5, CUDA events) for fp32, TF32 and bf16, plus a
    1 GiB device copy: the peaks the fp32 configurations can be read against."""
    n = 8192
    out = {}
    a32 = torch.randn(n, n, device=dev)
    b32 = torch.randn(n, n, device=dev)
    saved = torch.backends.cuda.matmul.allow_tf32

    def best(fn, reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts) * 1e-3
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        out["fp32_tflops"] = 2 * n ** 3 / best(lambda: torch.matmul(a32, b32), 3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = True
        out["tf32_tflops"] = 2 * n ** 3 / best(lambda: torch.matmul(a32, b32), 5) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved
    a16, b16 = a32.bfloat16(), b32.bfloat16()
    out["bf16_tflops"] = 2 * n ** 3 / best(lambda: torch.matmul(a16, b16), 5) / 1e12
    src = torch.empty(1 << 28, device=dev)
    dst = torch.empty_like(src)
    out["copy_gbs"] = 2 * src.numel() * 4 / best(lambda: dst.copy_(src), 5) / 1e9
    out["how"] = "torch.matmul 8192^3 best-of-N with CUDA events (fp32 = allow_tf32 off), 1 GiB fp32 copy_ (read+write)"
    return out


def stock_gpu_throughput(dev, B):
    """The reference's own nn.Module Python (oracle/_ref/refpy) over the reference's own kernels compiled for sm_100
    (oracle/_ref/libref_kernels.so) + cuDNN: what the stock libs/* build does on this GPU.  Timed like the reference
    times itself (synchronise + wall clock around model(x), datasets/scene_dataset.py:672-686), median of 5."""
    try:
        from oracle import refgpu, refpy
        if not (refpy.available() and refgpu.available()):
            return {"unavailable": "oracle/_ref (reference kernels + refpy) not built"}
        import util
        net0 = util.build_network("cpu")
        x = util.synthetic_batch(B, NPTS, start=0).unsqueeze(1).squeeze(1).to(dev)
        res = {}
        saved = torch.backends.cudnn.allow_tf32
        ref_net = refpy.reference_patchaugnet(net0.state_dict(), dev, "stock")
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    ref_net(x)
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    ref_net(x)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
            med = sorted(ts)[len(ts) // 2]
            res["cudnn_tf32" if tf32 else "fp32"] = dict(ms_per_batch=med * 1e3, submaps_per_s=B / med)
        torch.backends.cudnn.allow_tf32 = saved
        res["what"] = ("reference patch_aug_net.Network + pointops.py (unchanged Python) over the reference's CUDA kernels built "
                       "for sm_100 + PyTorch/cuDNN conv/BN/matmul; batch %d" % B)
        return res
    except Exception as ex:                                    # a baseline leg must never take the bench line down
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}


def pointnetvlad_cpu():
    """BASELINE.json configs[0] / SURVEY 8(d) CPU baseline (i): PointNetVLAD forward on the host cores (pure PyTorch in the
    reference; the reference's own file when oracle/_ref/refpy travelled, this repo's mirror of it otherwise)."""
    try:
        import util
        torch.set_num_threads(os.cpu_count() or 1)
        which = "mirror patchaugnet_b200.pointnet_vlad"
        net = None
        try:
            from oracle import refpy
            if refpy.available():
                import importlib.util as iu
                spec = iu.spec_from_file_location("ref_pointnetvlad", os.path.join(refpy.REFPY, "place_recognition", "pointnet_vlad", "PointNetVlad.py"))
                mod = iu.module_from_spec(spec)
                spec.loader.exec_module(mod)
                net = mod.PointNetVlad(global_feat=True, feature_transform=True, max_pool=False, output_dim=256, num_points=NPTS)
                which = "reference place_recognition/pointnet_vlad/PointNetVlad.py"
        except Exception:
            net = None
        if net is None:
            from patchaugnet_b200.pointnet_vlad import PointNetVlad
            net = PointNetVlad(num_points=NPTS, global_feat=True, feature_transform=True, max_pool=False, output_dim=256)
        net.eval()
        out = {}
        for b in (1, 8):
            x = util.synthetic_batch(b, NPTS, start=0)
            with torch.no_grad():
                net(x)
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    net(x)
                    ts.append(time.perf_counter() - t0)
            med = sorted(ts)[2]
            out[f"batch{b}"] = dict(ms_per_forward=med * 1e3, submaps_per_s=b / med)
        out.update(cores=os.cpu_count(), torch_threads=torch.get_num_threads(), module=which)
        return out
    except Exception as ex:
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}


def stock_pptnet(dev, ppt, x):
    """configs[2] denominator: the reference's pptnet.Network (unchanged Python) over the reference's kernels + cuDNN, same batch."""
    try:
        from oracle import refgpu, refpy
        if not (refpy.available() and refgpu.available()):
            return {"unavailable": "oracle/_ref not built"}
        ref_net = refpy.reference_pptnet(ppt.state_dict(), dev, "stock")
        saved = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            for _ in range(2):
                ref_net(x)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                ref_net(x)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
        torch.backends.cudnn.allow_tf32 = saved
        med = sorted(ts)[1]
        del ref_net
        torch.cuda.empty_cache()
        return dict(ms_per_batch=med * 1e3, submaps_per_s=x.shape[0] / med, what="reference pptnet.Network + pointops.py over the reference's "
                    "kernels built for sm_100 + cuDNN, fp32, batch %d" % x.shape[0])
    except Exception as ex:
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}


def stock_train_step(dev, net, feed, anchors):
    """configs[4] denominator: the same training step with the reference's own Network / pointops.py over the reference's kernels
    (fp32 atomics backward, cuDNN BatchNorm); the loss assembly is this repo's restatement of run_model (the reference's is broken
    as shipped) with this repo's chamfer kernels, so only the model forward/backward differs."""
    try:
        from oracle import refgpu, refpy
        from patchaugnet_b200 import training
        if not (refpy.available() and refgpu.available()):
            return {"unavailable": "oracle/_ref not built"}
        ref = refpy.use_backend("stock")
        ref_net = ref.patch_aug_net.Network(param=ref.cfg_patchaugnet, use_a2a_recon=True, use_l2_norm=True)
        ref_net.load_state_dict(net.state_dict())
        ref_net = ref_net.to(dev).train()
        step = training.TrainStep(ref_net, torch.optim.Adam(ref_net.parameters(), lr=5e-4), n_anchors=anchors)
        for _ in range(3):
            step(feed)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            step(feed)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        med = sorted(ts)[1]
        del ref_net, step
        torch.cuda.empty_cache()
        return dict(ms_per_step=med * 1e3, clouds_per_s=feed.shape[0] / med,
                    what="reference patch_aug_net.Network (train mode) + pointops.py over the reference's kernels + cuDNN, same tuples")
    except Exception as ex:
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}


def other_configs(dev, world, rank):
    """Sub-records for the BASELINE.json configurations that are not the bench line."""
    import torch.distributed as dist
    import util
    from patchaugnet_b200 import retrieval
    out = {}
    # configs[2]: PPT-Net, batch 64 x 4096: fp32 contract (bf16 hi/lo operands) and the bf16 mode, single forwards and throughput mode
    try:
        ppt = util.build_pptnet(dev)
        xs = [torch.cat([util.synthetic_batch(16, NPTS, start=16 * j)] * 4).to(dev) for j in range(2)]
        rec = {}
        for mode in ("f32", "bf16"):
            ppt.compute_dtype = mode
            eng = ppt.engine()
            seq = [xs[i & 1] for i in range(8)]
            with torch.no_grad():
                for _ in range(2):
                    ppt(xs[0], return_feat=False)
                eng.forward_stream(seq[:2])
                torch.cuda.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                for _ in range(5):
                    ppt(xs[0], return_feat=False)
                e1.record()
                eng.forward_stream(seq)
                e2.record()
                ref_desc = eng.forward_stream(seq)
                eng.forward_stream(seq, coalesce=128)                      # workspaces of the 128-cloud launch shape
                torch.cuda.synchronize()
                e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e3.record()
                co_desc = eng.forward_stream(seq, coalesce=128)
                e4.record()
                torch.cuda.synchronize()
            ms, ms_stream, ms_co = e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / len(seq), e3.elapsed_time(e4) / len(seq)
            rec[mode] = dict(ms_per_batch=ms, submaps_per_s_per_gpu=64 / (ms * 1e-3), ms_per_batch_stream=ms_stream,
                             submaps_per_s_per_gpu_stream=64 / (ms_stream * 1e-3), ms_per_batch_stream_coalesced=ms_co,
                             submaps_per_s_per_gpu_stream_coalesced=64 / (ms_co * 1e-3),
                             coalesced_bit_identical=bool(torch.equal(ref_desc, co_desc)))
        if rank == 0:
            rec["stock_gpu"] = stock_pptnet(dev, ppt, xs[0])
        rec["what"] = ("PPT-Net eval, batch 64 x 4096, fused engine, one GPU; f32 = the reference's fp32 contract (bf16 hi/lo tensor-core "
                       "operands), bf16 = plain bf16 operands for FP modules / NetVLAD / attention (SA modules keep hi/lo); stream = "
                       "geometry of batch i+1 under the dense kernels of batch i; stream_coalesced = two consecutive batches per launch sequence "
                       "(128 clouds), bit-identical descriptors")
        out["cfg3_pptnet_b64"] = rec
        del ppt, xs
    except Exception as ex:
        out["cfg3_pptnet_b64"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
    # configs[3]: 10k-submap database + 2k queries: sharded extraction, ONE all-gather, sharded top-k, Recall counters
    try:
        n_db, n_q = 10000, 2000
        net = util.build_network(dev)
        g = torch.Generator(device=dev).manual_seed(4321)         # same clouds on every rank; each rank reads its shard only
        clouds = (torch.rand(n_db + n_q, NPTS, 3, generator=g, device=dev) * 2 - 1) * 0.57
        with torch.no_grad():                                        # warm-up: kernels / lazily loaded modules of the retrieval ops
            wd = retrieval.extract_descriptors(net, clouds[:512 * world], batch_size=32, device=dev)   # incl. the CUDA graphs of the coalesced launch shape (captured from 4 groups on, ~1 s)
        retrieval.evaluate_recall(wd, wd[: 8 * world], [{i} for i in range(8 * world)], top_k=25)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, em, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        with torch.no_grad():      # untimed first pass: workspaces of the ragged tail shapes, allocator growth
            retrieval.extract_descriptor_sets(net, [clouds[:n_db], clouds[n_db:]], batch_size=32, device=dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        with torch.no_grad():      # database and queries as ONE pipelined sequence per rank, one all_gather per set
            db, qd = retrieval.extract_descriptor_sets(net, [clouds[:n_db], clouds[n_db:]], batch_size=32, device=dev)
            em.record()
        positives = retrieval.pad_positives([{i} for i in range(n_q)], device=dev)
        retrieval.evaluate_recall(db, qd, positives, top_k=25)     # untimed first call: kernel variants of this (k, shape) load lazily
        e1.record()
        res = retrieval.evaluate_recall(db, qd, positives, top_k=25)
        e2.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(em), e1.elapsed_time(e2), e0.elapsed_time(em)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ext, t_ret, t_db = t.tolist()
        out["cfg4_retrieval_10k"] = dict(extract_ms=t_ext, retrieval_ms=t_ret, submaps_per_s=(n_db + n_q) / (t_ext * 1e-3),
                                         queries_per_s=n_q / (t_ret * 1e-3), k=res["k"], n_gpus=world,
                                         what="10000 db + 2000 query clouds resident in HBM, sharded by rank, one all_gather per "
                                              "descriptor set, brute-force top-k per query shard, hit counters all_reduced")
        del clouds, db, qd
    except Exception as ex:
        out["cfg4_retrieval_10k"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
    # configs[4]: training step, 16 anchors x 18 clouds x 4096 points per GPU (128 anchors over 8 GPUs), quadruplet + patch chamfer
    try:
        from patchaugnet_b200 import training
        torch.cuda.empty_cache()
        anchors = 16
        net = util.build_network(dev).train()
        model = training.build_ddp(net, dev)
        opt = torch.optim.Adam(model.parameters(), lr=5e-4)
        step = training.TrainStep(model, opt, n_anchors=anchors)
        g = torch.Generator(device=dev).manual_seed(77 + rank)
        feed = ((torch.rand(anchors * training.CLOUDS_PER_ANCHOR, 1, NPTS, 3, generator=g, device=dev) * 2 - 1) * 0.57)
        for _ in range(5):                                         # allocator growth, cuDNN algorithm choice, Adam state, DDP buckets
            step(feed)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        per_step = []
        for _ in range(7):                                         # median of 7 individually timed steps (max over ranks each)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss, _terms = step(feed)
            e1.record()
            torch.cuda.synchronize()
            per_step.append(e0.elapsed_time(e1))
        t = torch.tensor(per_step, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.median())
        rec = dict(ms_per_step=ms, anchors_per_gpu=anchors, clouds_per_gpu=anchors * training.CLOUDS_PER_ANCHOR, n_gpus=world,
                   clouds_per_s=world * anchors * training.CLOUDS_PER_ANCHOR / (ms * 1e-3), loss=float(loss),
                   grad_allreduce_bytes=training.grad_bytes(net), peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                   what="train-mode forward (batch statistics: cuDNN 1x1 convolutions + this repo's fused BatchNorm+ReLU kernels, pointops / "
                        "chamfer kernels with deterministic backward) + quadruplet + patch-chamfer loss + backward + Adam; "
                        "DistributedDataParallel over anchors")
        if world > 1:      # the one collective of the step, timed alone: a flat all-reduce of the gradient bytes
            flat = torch.empty(rec["grad_allreduce_bytes"] // 4, device=dev)
            dist.all_reduce(flat)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(5):
                dist.all_reduce(flat)
            a1.record()
            torch.cuda.synchronize()
            rec["allreduce_alone_ms"] = a0.elapsed_time(a1) / 5
            rec["allreduce_share_of_step"] = rec["allreduce_alone_ms"] / ms
        if world == 1:      # the same step replayed as one CUDA graph (single process): GPU-bound whatever the host does
            try:
                net_g = util.build_network(dev).train()
                gstep = training.GraphedTrainStep(net_g, torch.optim.Adam(net_g.parameters(), lr=5e-4, capturable=True), n_anchors=anchors)
                gstep(feed)
                gstep(feed)
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(5):
                    gloss, _t = gstep(feed)
                g1.record()
                torch.cuda.synchronize()
                rec["cuda_graph"] = dict(ms_per_step=g0.elapsed_time(g1) / 5, clouds_per_s=anchors * training.CLOUDS_PER_ANCHOR / (g0.elapsed_time(g1) / 5 * 1e-3),
                                         loss=float(gloss), what="training.GraphedTrainStep: forward + losses + backward + Adam captured once, replayed")
                del net_g, gstep
                torch.cuda.empty_cache()
            except Exception as ex:
                rec["cuda_graph"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
        if rank == 0 and world == 1:
            rec["stock_gpu"] = stock_train_step(dev, net, feed, anchors)
        out["cfg5_train_step"] = rec
        del net, model, opt, step, feed
        torch.cuda.empty_cache()
    except Exception as ex:
        out["cfg5_train_step"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip stock_gpu / pointnetvlad_cpu / measured_peaks / configs sub-records")
    ap.add_argument("--repeats", type=int, default=0, help="repetitions of the K-step timed region (0 = enough for >= 200 steps)")
    ap.add_argument("--mode", default="stream", choices=["stream", "graph", "eager"],
                    help="stream: 2-stream pipelined throughput mode (default); graph: one CUDA graph per step; eager")
    ap.add_argument("--dense-streams", type=int, default=2)
    ap.add_argument("--stream-graphs", type=int, default=1, help="stream mode: replay the geometry / dense launch sequences as CUDA graphs")
    ap.add_argument("--tc-ctas", type=int, default=0, help="cap on persistent tensor-core CTAs (0 = one per SM)")
    ap.add_argument("--tc-tune", type=int, default=1, help="pab_tune_tensor_core bits (1 on, +4 CTA-pair multicast, +8 static tiles)")
    ap.add_argument("--fp-order", type=int, default=1, help="FP modules walk their points in Morton order (0 = index order)")
    ap.add_argument("--fps-threads", type=int, default=0, help="force the FPS CTA size (0 = automatic)")
    ap.add_argument("--fps-cpc", type=int, default=1, help="clouds per FPS CTA in stream mode (1..3)")
    ap.add_argument("--reserve-fps-sms", type=int, default=0, help="stream mode: cap the persistent tensor-core kernels at (SMs - FPS CTAs)")
    ap.add_argument("--dynamic-tiles", type=int, default=1, help="stream mode: tensor-core CTAs draw tiles from a counter")
    ap.add_argument("--coalesce", type=int, default=0, help="stream mode: clouds per launch sequence of the timed region (0 = one sequence per "
                    "batch, the configuration named in config.global_batch); the `coalesced` record and `e2e` use 128")
    ap.add_argument("--slots", type=int, default=4, help="stream mode: workspace slots of the three-stage pipeline")
    ap.add_argument("--fps-stream", type=int, default=1, help="stream mode: first-level FPS on a stream of its own (three-stage pipeline)")
    ap.add_argument("--fps-pruned", type=int, default=0, help="1 = pruned FPS sampler (exact, but slower at these sizes)")
    ap.add_argument("--prio", default="0,0", help="CUDA stream priorities geometry,dense (lower = higher priority)")
    ap.add_argument("--graph", type=int, default=None, help="deprecated alias: 1 -> --mode graph, 0 -> --mode eager")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.graph is not None:
        args.mode = "graph" if args.graph else "eager"

    import torch.distributed as dist
    import util
    from patchaugnet_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K, B = max(3, args.warmup), args.steps, args.batch

    net = util.build_network(dev)                        # random-init weights of the reference architecture (tests/util.py)
    eng = net.engine()
    if args.tc_ctas:
        eng.reserve_fps_sms = False
        L.lib().pab_tune_tc_max_ctas(args.tc_ctas)
    eng.fps_clouds_per_cta = args.fps_cpc
    L.lib().pab_tune_tensor_core(args.tc_tune)
    eng.tc_tune = args.tc_tune
    L.lib().pab_tune_fps_threads(args.fps_threads)
    L.lib().pab_tune_fps_pruned(args.fps_pruned)
    eng.dense_streams = max(1, min(3, args.dense_streams))
    eng.stream_graphs = bool(args.stream_graphs)
    eng.fps_stream = bool(args.fps_stream)
    eng.stream_slots = args.slots
    eng.reserve_fps_sms = bool(args.reserve_fps_sms)
    eng.stream_dynamic_tiles = bool(args.dynamic_tiles)
    eng.fp_row_order = bool(args.fp_order)
    eng.stream_priorities = tuple(int(v) for v in args.prio.split(","))
    lib = L.lib()

    # input pool larger than L2 (126 MB): 104 batches x 1.5 MB = 164 MB, distinct seeded clouds per rank
    pool_batches = 104
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    base = torch.rand(pool_batches * B, NPTS, 3, generator=g) * 2 - 1
    base = base - base.mean(1, keepdim=True)
    base = base / base.norm(dim=2).max(dim=1)[0][:, None, None]
    pool = base.to(dev).view(pool_batches, B, NPTS, 3)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up + per-stage profile (eager, events around every stage) -----------------------------------------
    if os.environ.get("PAB_SN"):                       # experiment hook: "enable,ctas_per_sm" of pab_tune_sa_narrow
        lib.pab_tune_sa_narrow(*[int(v) for v in os.environ["PAB_SN"].split(",")])
    eng.enable_stage_timing()
    with torch.no_grad():
        for i in range(W):
            eng(pool[i % pool_batches], clone=False)
    torch.cuda.synchronize()
    st = eng.stage_times_ms()
    eng.disable_stage_timing()
    stage_ms = {k: float(np.median(v[1:] if len(v) > 1 else v)) for k, v in st.items()}   # median: one slow launch must not pick the dominant kernel
    work = eng.stage_work(B, NPTS)

    def occupancy(stage):
        # FPS runs one CTA per cloud (B of the 148 SMs) and, in stream mode, off the critical path next to the dense
        # kernels; every other stage fills the machine.  "Dominant" = largest share of SM-time, not of wall time.
        return min(1.0, B / 148.0) if stage.startswith("fps") else 1.0
    dominant = max(stage_ms, key=lambda k: stage_ms[k] * occupancy(k))

    mode = args.mode
    if mode == "graph":
        eng.capture_graph(B, NPTS)
        with torch.no_grad():
            for i in range(2):
                eng(pool[i], clone=False, return_feat=False)
    elif mode == "stream":
        with torch.no_grad():
            eng.forward_stream([pool[i] for i in range(4)])          # creates the side streams / second workspace
    gathered = torch.empty(world * K * B, 256, device=dev) if world > 1 else None
    local_desc = torch.empty(K * B, 256, device=dev)

    # ---- timed region: exactly K steps, device-resident inputs; the region is repeated R times (each repetition is
    # bracketed by barrier + synchronise, timed with CUDA events, max over ranks) and the MEDIAN repetition is reported,
    # so that the measurement spans >= 200 steps / several clock samples even when K is small -------------------------
    R = args.repeats if args.repeats > 0 else max(1, -(-200 // K))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    region_ms = []
    launches = 0
    for rep in range(R):
        sync_all()
        lib.pab_reset_launch_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            if mode == "stream":     # throughput mode: batch i+1's geometry overlaps batch i's dense kernels (2 streams)
                eng.forward_stream([pool[(W + rep * K + i) % pool_batches] for i in range(K)], out=local_desc, coalesce=args.coalesce)
            else:
                for i in range(K):
                    desc = eng(pool[(W + rep * K + i) % pool_batches], clone=False, return_feat=False)
                    local_desc[i * B:(i + 1) * B].copy_(desc)
            if world > 1:   # the one collective of the path: all-gather of the 256-D descriptors before retrieval
                dist.all_gather_into_tensor(gathered, local_desc)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        region_ms.append(ms)
        launches = lib.pab_num_launches()
    if mode == "graph" or (mode == "stream" and getattr(eng, "last_stream_used_graphs", False)):
        launches = K * eng.launches_per_forward()        # graph replays do not pass through the C ABI counter
    elapsed_ms = float(np.median(region_ms))
    value = world * B * K / (elapsed_ms * 1e-3)
    # the same device-resident region with consecutive batches concatenated into launch sequences of 128 clouds (what the public API
    # retrieval.extract_descriptors does): bit-identical descriptors, reported next to the per-batch number, never instead of it
    coalesced = None
    if mode == "stream" and not args.coalesce:
        Kc = max(K, 64) // 4 * 4
        co_desc = torch.empty(Kc * B, 256, device=dev)
        co_ms = []
        with torch.no_grad():
            eng.forward_stream([pool[i % pool_batches] for i in range(8)], coalesce=4 * B)          # workspaces / graphs of this shape
            for rep in range(3):
                sync_all()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.forward_stream([pool[(rep * Kc + i) % pool_batches] for i in range(Kc)], out=co_desc, coalesce=4 * B)
                e1.record()
                sync_all()
                ms = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ms], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = t.item()
                co_ms.append(ms)
        cms = float(np.median(co_ms))
        coalesced = dict(value=world * B * Kc / (cms * 1e-3), unit=UNIT, clouds_per_launch_sequence=4 * B, steps=Kc, ms_per_step=cms / Kc,
                         what="forward_stream(coalesce=128): four consecutive 32-cloud batches per launch sequence, descriptors bit-identical "
                              "to the per-batch sequences (tests/test_model_gpu.py)")
        del co_desc
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel: events around that stage only, K eager steps over the same rotating inputs -----------
    eng._graphs.clear()
    eng.enable_stage_timing([dominant])
    with torch.no_grad():
        for i in range(K):
            eng(pool[(W + i) % pool_batches], clone=False)
    torch.cuda.synchronize()
    dom_ms = float(np.mean(eng.stage_times_ms()[dominant]))
    eng.disable_stage_timing()
    pk = peaks()
    clock_hz = 1.965e9

    def roof_entry(stage, ms):
        wk = work[stage]
        if wk["flops"] > 0:
            ach = wk["flops"] / (ms * 1e-3) / 1e12
            return dict(kernel=stage, bound="tensor", achieved=ach, peak=pk["tf_sustained"], unit="TFLOP/s", frac=ach / pk["tf_sustained"],
                        ms_per_launch=ms, algorithmic_flops_per_launch=wk["flops"], algorithmic_bytes_per_launch=wk["bytes"])
        ach = wk["bytes"] / (ms * 1e-3) / 1e9
        e = dict(kernel=stage, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], ms_per_launch=ms,
                 algorithmic_bytes_per_launch=wk["bytes"])
        if "units" in wk:   # scan kernels are issue/latency bound: also report against the fp32 issue bound
            per_unit = 9.0                      # sub x3, mul, fma x2, min/compare, select, index bookkeeping
            sms = min(148, B) if stage.startswith("fps") else 148
            bound = sms * 128 * clock_hz / per_unit
            e["issue_bound"] = dict(units=wk["units"], unit=wk["unit"], achieved_per_s=wk["units"] / (ms * 1e-3),
                                    bound_per_s=bound, frac=wk["units"] / (ms * 1e-3) / bound, sms_used=sms)
        return e

    # stages that run on the tensor-core kernels issue 3 bf16 MMAs per algorithmic product
    tc_stages = {"sa0", "sa1", "sa2", "fp0", "fp1", "vlad0", "vlad1", "vlad2"}

    def traffic_of(stage):
        """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/r0N_traffic.json), B=32 only."""
        try:
            name = "r02_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_traffic.json")) else "r01_traffic.json"
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)["dram_bytes_per_launch"]
            return int(t[stage]) if B == 32 and stage in t else None
        except (OSError, KeyError, ValueError):
            return None

    roof = roof_entry(dominant, dom_ms)
    roof["traffic"] = traffic_of(dominant)
    if roof["bound"] == "tensor" and dominant in tc_stages:
        roof["issued"] = dict(mma_flops_per_launch=3 * work[dominant]["flops"], tensor_pipe_frac=3 * roof["frac"],
                              note="bf16 hi/lo split: hi*hi + lo*hi + hi*lo per product")
    roof["share_of_step"] = dom_ms / sum(stage_ms.values())
    roof["peak_source"] = (f"{pk['src']} bf16 dense, sustained (kernel timed inside the step); algorithmic fp32-equivalent FLOPs — the "
                           "tcgen05 path issues 3 bf16 MMAs per product") if roof["bound"] == "tensor" else f"{pk['src']} copy bandwidth"
    roof_all = [roof_entry(k, v) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1]) if v > 0.02]

    # ---- end to end through the public API: pinned host clouds -> extract_descriptors(net, ...) -> host descriptors ----
    from patchaugnet_b200 import retrieval
    # weak scaling: every rank extracts K*B clouds of a common (world*K*B)-cloud database, then one all-gather
    g2 = torch.Generator(device="cpu").manual_seed(999)
    host_clouds = (torch.rand(world * K * B, NPTS, 3, generator=g2) * 2 - 1).mul_(0.57).contiguous().pin_memory()
    out_host = torch.empty(world * K * B, 256).pin_memory()
    with torch.no_grad():
        retrieval.extract_descriptors(net, host_clouds[: world * 16 * B], batch_size=B, device=dev)    # warm-up (incl. the graphs of the coalesced shape)
    e2e_runs = []
    for _ in range(3):                                      # median of 3 repetitions of the K-step region
        sync_all()
        t0 = time.perf_counter()
        with torch.no_grad():
            d = retrieval.extract_descriptors(net, host_clouds, batch_size=B, device=dev)   # H2D + K steps per rank + all-gather
            out_host.copy_(d)                                                               # D2H of the result
        sync_all()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = sorted(e2e_runs)[1]
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e = dict(value=world * K * B / e2e_s, unit=UNIT, h2d_bytes_per_step=B * NPTS * 3 * 4, d2h_bytes_per_step=world * B * 256 * 4,
               api="patchaugnet_b200.retrieval.extract_descriptors(net, pinned_host_clouds, batch_size=32) + .cpu()",
               clouds_per_launch_sequence=retrieval.LAUNCH_BATCH,
               note="uploads in batches of 32; the engine concatenates four consecutive batches per launch sequence (bit-identical descriptors)")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        v, done, el = cpu_oracle_throughput()
        cpu = dict(value=v, unit=UNIT, cores=os.cpu_count(), kind="port",
                   sample=f"{done} clouds x {NPTS} pts in {el:.1f} s (32 per call), oracle PatchAugNet eval forward fp32 "
                          f"(C scans OpenMP over {os.cpu_count()} threads, dense layers torch-CPU)")

    extras = {}
    if not args.no_extras:
        if rank == 0 and world == 1:
            extras["stock_gpu"] = stock_gpu_throughput(dev, B)
            extras["measured_peaks"] = measure_matmul_peaks(dev)
            if not args.no_cpu_baseline:
                extras["pointnetvlad_cpu"] = pointnetvlad_cpu()
        extras["configs"] = other_configs(dev, world, rank)          # collective at world > 1: every rank takes part

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"PatchAugNet descriptor extraction, batch {B} x {NPTS}-pt synthetic clouds per GPU, fp32, eval "
                                   "(BASELINE.json configs[1])", "global_batch": world * B, "l2": "inputs_larger_than_l2 (164 MB rotating pool)",
                       "launch": mode, "parallelism": f"dp{world}, shard-by-submap, one all_gather of descriptors"},
            "roofline": roof, "roofline_all": roof_all, "cpu_baseline": cpu, "e2e": e2e, "coalesced": coalesced, "gpu_launches": int(launches), "clocks": clocks,
            "timed_region": {"repeats": R, "steps_per_repeat": K, "ms_per_repeat": [round(v, 4) for v in region_ms],
                             "reported": "median repetition"},
            "stage_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}, **extras}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
