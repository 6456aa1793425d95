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


def visit(p, v, npts=4096):
    import util
    return util.place_visit(p, v, npts, rotate=ROTATE)


ROTATE = False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", type=int, default=10000)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--top-k", type=int, default=25)
    ap.add_argument("--rotate", action="store_true", help="random yaw per visit (random-init weights are not rotation invariant)")
    ap.add_argument("--oracle-check", type=int, default=0, metavar="Q",
                    help="also run the CPU oracle forward on the first 4Q database clouds and Q queries and compare Recall@N")
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
        # shard + one pipelined sequence per rank + one all-gather per set
        db, qd = retrieval.extract_descriptor_sets(net, [db_clouds, q_clouds], batch_size=args.batch, device=dev)
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
                   weights="random-init + calibrated BatchNorm statistics (tests/util.fill_state_dict): recall measures pipeline "
                           "agreement, not a trained model")
        if args.oracle_check and world == 1:      # single-rank only: evaluate_recall is collective, clouds are sharded
            from oracle import model                     # the checker, on a bounded sub-database (CPU: ~7 clouds/s)
            nq_o, ndb_o = args.oracle_check, min(args.db, 4 * args.oracle_check)
            perms = [np.arange(20)] * 3
            fwd = lambda x: torch.cat([model.patchaugnet_forward(net.state_dict(), util.PATCHAUGNET_CFG, x[i:i + 16].numpy(),
                                                                 perms=perms)["desc"] for i in range(0, len(x), 16)])
            o_db, o_q = fwd(db_clouds[:ndb_o]), fwd(q_clouds[:nq_o])
            r_new = retrieval.evaluate_recall(db[:ndb_o], qd[:nq_o], positives[:nq_o], top_k=args.top_k)
            r_ref = retrieval.evaluate_recall(o_db.to(dev), o_q.to(dev), positives[:nq_o], top_k=args.top_k)
            out["oracle_check"] = dict(db=ndb_o, queries=nq_o,
                                       max_abs_desc_diff=float(max((db[:ndb_o].cpu() - o_db).abs().max(), (qd[:nq_o].cpu() - o_q).abs().max())),
                                       recall_new=[float(r_new["recall"][i]) for i in (0, 4, 9)],
                                       recall_oracle=[float(r_ref["recall"][i]) for i in (0, 4, 9)],
                                       recall_identical=bool(np.array_equal(r_new["recall"], r_ref["recall"])))
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
