#!/usr/bin/env python
"""BASELINE.json configs[3]: PatchAugNet retrieval over a synthetic submap database, descriptors sharded over the
ranks + one NCCL all-gather, Recall@1/5/10 and top-1% recall with the reference's first-hit rule.

    python scripts/retrieval_eval.py [--db 10000] [--queries 2000]                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/retrieval_eval.py --db 10000 --queries 2000                              # N GPUs

Synthetic places: place p = seeded base cloud of 8 random planes + 4 random cylinders + 10 %
uniform noise (8192 points); a visit = base -> random yaw rotation (utils/loading_pointclouds.py:102-128) -> jitter
N(0, 0.005) clipped at 0.05 (:163-174) -> random 4096-subset -> unit-ball normalisation (:51-63).  Database = one visit
per place, queries = second visits of the first Q places; positive <=> same place id.

Checks printed in the JSON line: GPU brute-force top-k vs sklearn KDTree on the same descriptors (distances within
1e-4, identical Recall@N), and — when run on several ranks — that every rank holds the same gathered database.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def base_place(p, n=8192):
    rng = np.random.default_rng(10_000 + p)
    pts = []
    per = int(n * 0.9) // 12
    for _ in range(8):                                   # planes: random point + two in-plane axes
        o = rng.uniform(-1, 1, 3)
        a, b = rng.normal(size=3), rng.normal(size=3)
        a /= np.linalg.norm(a); b -= a * (a @ b); b /= np.linalg.norm(b)
        uv = rng.uniform(-0.8, 0.8, (per, 2))
        pts.append(o + uv[:, :1] * a + uv[:, 1:] * b)
    for _ in range(4):                                   # vertical-ish cylinders
        o = rng.uniform(-1, 1, 3)
        r = rng.uniform(0.05, 0.3)
        th = rng.uniform(0, 2 * np.pi, per)
        h = rng.uniform(-0.8, 0.8, per)
        pts.append(o + np.stack([r * np.cos(th), r * np.sin(th), h], 1))
    pts = np.concatenate(pts)
    noise = rng.uniform(-1.5, 1.5, (n - len(pts), 3))
    return np.concatenate([pts, noise]).astype(np.float32)


ROTATE = False


def visit(p, v, npts=4096):
    rng = np.random.default_rng(1_000_000 * (v + 1) + p)
    base = base_place(p)
    ang = rng.uniform(0, 2 * np.pi) if ROTATE else 0.0
    c, s = np.cos(ang), np.sin(ang)
    rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)          # rotate_point_cloud: about the up axis
    pts = base @ rot
    pts = pts + np.clip(0.005 * rng.normal(size=pts.shape), -0.05, 0.05).astype(np.float32)
    pts = pts[rng.choice(len(pts), npts, replace=False)]
    pts = pts - pts.mean(0, keepdims=True)                                    # normalize_point_cloud
    return (pts / np.max(np.linalg.norm(pts, axis=1))).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", type=int, default=10000)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--top-k", type=int, default=25)
    ap.add_argument("--rotate", action="store_true", help="random yaw per visit (random-init weights are not rotation invariant)")
    args = ap.parse_args()
    global ROTATE
    ROTATE = args.rotate

    import util
    from patchaugnet_b200 import retrieval

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.perf_counter()
    # every rank only synthesises its own shard of the clouds (the rest stays zeros and is never read)
    lo, hi = retrieval.shard_range(args.db, rank, world)
    db_clouds = torch.zeros(args.db, 4096, 3).pin_memory()
    for p in range(lo, hi):
        db_clouds[p] = torch.from_numpy(visit(p, 0))
    qlo, qhi = retrieval.shard_range(args.queries, rank, world)
    q_clouds = torch.zeros(args.queries, 4096, 3).pin_memory()
    for p in range(qlo, qhi):
        q_clouds[p] = torch.from_numpy(visit(p, 1))
    t_gen = time.perf_counter() - t0

    net = util.build_network(dev)
    with torch.no_grad():
        retrieval.extract_descriptors(net, db_clouds[lo:lo + 2 * args.batch], batch_size=args.batch, device=dev)   # warm-up (local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    with torch.no_grad():
        db = retrieval.extract_descriptors(net, db_clouds, batch_size=args.batch, device=dev)       # shard + all-gather
        qd = retrieval.extract_descriptors(net, q_clouds, batch_size=args.batch, device=dev)
    e1.record()
    positives = [{i} for i in range(args.queries)]
    res = retrieval.evaluate_recall(db, qd, positives, top_k=args.top_k)
    e2.record()
    torch.cuda.synchronize()
    t_extract, t_retr = e0.elapsed_time(e1), e1.elapsed_time(e2)
    if world > 1:
        t = torch.tensor([t_extract, t_retr], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_extract, t_retr = t.tolist()
        chk = torch.tensor([db.double().sum().item(), qd.double().sum().item()], device=dev, dtype=torch.float64)
        lo_chk, hi_chk = chk.clone(), chk.clone()
        dist.all_reduce(lo_chk, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_chk, op=dist.ReduceOp.MAX)
        same_everywhere = bool(torch.equal(lo_chk, hi_chk))
    else:
        same_everywhere = True

    out = None
    if rank == 0:
        from sklearn.neighbors import KDTree
        k, thr = retrieval.real_top_k(args.db, args.top_k)
        k = min(k, args.db)
        nq_chk = min(args.queries, 500)
        kd, ki = KDTree(db.cpu().numpy()).query(qd[:nq_chk].cpu().numpy(), k=k)      # place_recognition_dataset.py:60
        gd, gi = retrieval.retrieval_topk(db, qd[:nq_chk], k)
        hits, one_pct, ev = retrieval.recall_counts(ki, positives[:nq_chk], args.top_k, thr)
        hits_g, one_g, ev_g = retrieval.recall_counts(gi.cpu().numpy(), positives[:nq_chk], args.top_k, thr)
        out = dict(config="PatchAugNet retrieval, synthetic places", db=args.db, queries=args.queries, n_gpus=world,
                   recall_at_1=float(res["recall"][0]), recall_at_5=float(res["recall"][4]), recall_at_10=float(res["recall"][9]),
                   one_percent_recall=float(res["one_percent_recall"]), evaluated=res["evaluated"], k=res["k"],
                   extract_ms=t_extract, retrieval_ms=t_retr, submaps_per_s=(args.db + args.queries) / (t_extract * 1e-3),
                   kdtree_max_abs_dist_diff=float(np.abs(gd.cpu().numpy() - kd).max()),
                   kdtree_recall_identical=bool(np.array_equal(hits, hits_g) and one_pct == one_g),
                   gathered_db_identical_on_all_ranks=same_everywhere, data_gen_s=t_gen,
                   weights="random-init (tests/util.fill_state_dict): recall measures pipeline agreement, not a trained model")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
