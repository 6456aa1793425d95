"""Database descriptor extraction, sharded over ranks, and GPU retrieval + Recall@N.

Replaces, for the synthetic-data harness, the reference's evaluation flow
  ``SceneDataSet.make_descs``           datasets/scene_dataset.py:494-711   (batched model forward, one GPU)
  ``KDTree(db).query(q, k)``            datasets/place_recognition_dataset.py:60, scene_dataset.py:1052 (CPU, per query)
  ``SceneDataSet.get_recall_precision`` datasets/scene_dataset.py:1016-1099 (first-hit cumsum recall, top-1 % recall)

Multi-GPU: submaps are independent units, so the database is partitioned by contiguous index
range over ranks (one process per GPU, weights replicated), each rank extracts its shard, and ONE
``all_gather_into_tensor`` of the (N/W, 256) fp32 descriptors over NCCL/NVLink gives every rank the full database.
Queries are sharded the same way; each rank answers its queries with the brute-force top-k kernel
(``pab_retrieval_topk``) and the per-rank hit counters are summed with one small ``all_reduce``.
Results are identical for every world size.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`: the first n % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


_COPY_STREAMS = {}


def _copy_stream(device):
    key = torch.device(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=key)
    return _COPY_STREAMS[key]


LAUNCH_BATCH = 128     # clouds per launch sequence of the fused engine in throughput mode (engine.forward_stream(coalesce=...))


def extract_descriptors(extract_fn, clouds, batch_size=32, device=None, dim=256, group=None, out_device=None,
                        super_chunk=64, launch_batch=None):
    """Descriptors of this rank's shard of ``clouds`` (M,N,3), all-gathered so every rank returns the full (M, dim).

    ``extract_fn`` is either a ``patchaugnet_b200.patch_aug_net.Network`` / ``pptnet.Network`` in eval mode on CUDA — then the shard runs
    through the fused engine in throughput mode (``FusedPatchAugNet.forward_stream``: geometry of batch i+1 overlapped
    with the dense kernels of batch i) — or any callable ``x (b,1,N,3) on device -> (b,dim)``.
    ``clouds`` may live on the host (pinned memory recommended: the copies are issued non-blocking, ``super_chunk``
    batches at a time) or on the device.  ``batch_size`` is the granularity of the uploads; the fused engine concatenates
    consecutive batches into launch sequences of up to ``launch_batch`` clouds (default ``LAUNCH_BATCH`` = 128; descriptors are
    bit-identical for any value, ``launch_batch=batch_size`` keeps one sequence per batch).
    """
    return extract_descriptor_sets(extract_fn, [clouds], batch_size, device, dim, group, out_device, super_chunk, launch_batch)[0]


def extract_descriptor_sets(extract_fn, cloud_sets, batch_size=32, device=None, dim=256, group=None, out_device=None,
                            super_chunk=64, launch_batch=None):
    """``extract_descriptors`` for several cloud sets at once (database and queries of an evaluation): every set is sharded over
    the ranks by contiguous index range, this rank's shards of ALL sets run through ONE pipelined batch sequence (the
    pipeline fills and drains once, not once per set), then each set gets its own all_gather.  Returns a list of (M_i, dim)."""
    rank, world = _dist_info(group)
    first = cloud_sets[0]
    device = device if device is not None else (first.device if first.is_cuda else torch.device("cpu"))
    shards = [shard_range(c.shape[0], rank, world) for c in cloud_sets]
    locals_ = [torch.empty(hi - lo, dim, dtype=torch.float32, device=device) for lo, hi in shards]
    engine = None
    if hasattr(extract_fn, "engine") and hasattr(extract_fn, "fusable") and torch.device(device).type == "cuda" \
            and not extract_fn.training and extract_fn.fusable():
        engine = extract_fn.engine()
    if engine is not None:
        # jobs: (set, start, end) batches of this rank; full batches of every set form one pipelined sequence, ragged tails follow
        full_jobs, tail_jobs = [], []
        for si, (lo, hi) in enumerate(shards):
            for s0 in range(lo, hi, batch_size):
                e0 = min(hi, s0 + batch_size)
                (full_jobs if e0 - s0 == batch_size else tail_jobs).append((si, s0, e0))
        host_side = not first.is_cuda
        copy_stream = _copy_stream(device) if host_side else None

        def stage(jobs):
            """device batches of `jobs` (+ per-batch upload events when the clouds live on the host)"""
            if not host_side:
                return [cloud_sets[si][s0:e0] for si, s0, e0 in jobs], None
            # one upload per batch on a copy stream, each followed by an event: batch i's kernels wait for ITS upload only,
            # so the copies of later batches run under the compute of earlier ones (the make_descs loop of the reference
            # uploads and computes strictly in turn, scene_dataset.py:672-686)
            batches, events = [], []
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                for si, s0, e0 in jobs:
                    batches.append(cloud_sets[si][s0:e0].to(device, non_blocking=True))
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    events.append(ev)
            for bt in batches:
                bt.record_stream(torch.cuda.current_stream())
            return batches, events

        for c0 in range(0, len(full_jobs), super_chunk):
            jobs = full_jobs[c0:c0 + super_chunk]
            batches, events = stage(jobs)
            out = engine.forward_stream(batches, ready_events=events,
                                        coalesce=LAUNCH_BATCH if launch_batch is None else launch_batch)
            for j, (si, s0, e0) in enumerate(jobs):
                locals_[si][s0 - shards[si][0]:e0 - shards[si][0]] = out[j * batch_size:(j + 1) * batch_size]
        if tail_jobs:
            batches, events = stage(tail_jobs)
            for j, (si, s0, e0) in enumerate(tail_jobs):
                if events is not None:
                    torch.cuda.current_stream().wait_event(events[j])
                takes_clone = "clone" in engine.forward.__code__.co_varnames        # FusedPatchAugNet: return the workspace view
                locals_[si][s0 - shards[si][0]:e0 - shards[si][0]] = (
                    engine(batches[j], return_feat=False, clone=False) if takes_clone else engine(batches[j], return_feat=False))
    else:
        for si, (lo, hi) in enumerate(shards):
            for s0 in range(lo, hi, batch_size):
                e0 = min(hi, s0 + batch_size)
                x = cloud_sets[si][s0:e0].to(device, non_blocking=True).unsqueeze(1)
                out = extract_fn(x)
                locals_[si][s0 - lo:e0 - lo] = out[0] if isinstance(out, (tuple, list)) else out
    results = []
    for si, (lo, hi) in enumerate(shards):
        local, M = locals_[si], cloud_sets[si].shape[0]
        if world == 1:
            results.append(local if out_device is None else local.to(out_device))
            continue
        # equal-size shards are required by all_gather_into_tensor: pad to the largest shard, trim after
        per = (M + world - 1) // world
        send = torch.zeros(per, dim, dtype=torch.float32, device=device)
        send[: hi - lo] = local
        recv = torch.empty(world * per, dim, dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(recv, send, group=group)
        parts = []
        for r in range(world):
            rlo, rhi = shard_range(M, r, world)
            parts.append(recv[r * per: r * per + (rhi - rlo)])
        full = torch.cat(parts, 0)
        results.append(full if out_device is None else full.to(out_device))
    return results


def retrieval_topk(db, queries, k):
    """Exact k nearest database descriptors for every query on the GPU (ties to the lower index).
    db (Ndb,D), queries (Nq,D) float32 CUDA -> dist (Nq,k) float32 ascending Euclidean, ind (Nq,k) int32."""
    L.require_cuda(db, queries)
    db = db.contiguous().float()
    queries = queries.contiguous().float()
    ndb, d = db.shape
    nq = queries.shape[0]
    out_d = torch.empty(nq, k, dtype=torch.float32, device=db.device)
    out_i = torch.empty(nq, k, dtype=torch.int32, device=db.device)
    if k <= 128 and nq > 0:
        # database split over the grid + merge: the same exact result, and a small query shard still fills the machine
        ws = torch.empty(L.lib().pab_retrieval_topk_workspace_bytes(nq, k), dtype=torch.uint8, device=db.device)
        L.check(L.lib().pab_retrieval_topk_split(L.ptr(db), ndb, L.ptr(queries), nq, d, k, L.ptr(out_d), L.ptr(out_i), L.ptr(ws),
                                                 L.stream_ptr()), "retrieval_topk_split")
    else:
        L.check(L.lib().pab_retrieval_topk(L.ptr(db), ndb, L.ptr(queries), nq, d, k, L.ptr(out_d), L.ptr(out_i), L.stream_ptr()),
                "retrieval_topk")
    return out_d, out_i


def real_top_k(n_db, top_k=25):
    """k actually queried: max(top_k + 1, round(n_db/100) + 1)  (scene_dataset.py:1026-1029)."""
    threshold = max(int(round(n_db / 100.0)), 1)
    return max(top_k + 1, threshold + 1), threshold


def recall_counts(ind, positives, top_k, threshold, query_db_index=None):
    """First-hit counters of scene_dataset.py:1056-1081 for one shard of queries.

    ind (Nq,K) retrieved database indices (ascending distance); positives: list of sets of database indices;
    query_db_index: the query's own database index when the query set is part of the database (its self-match is
    skipped like `add_one_more`, :1053-1055).  Returns (recall_hits[top_k], one_percent_hits, evaluated).
    """
    ind = np.asarray(ind)
    hits = np.zeros(top_k, dtype=np.int64)
    one_pct = 0
    evaluated = 0
    for qi in range(ind.shape[0]):
        pos = positives[qi]
        if not pos:
            continue                                   # :1044-1045
        evaluated += 1
        row = ind[qi]
        if query_db_index is not None:
            row = row[1:]                              # the first neighbour is the query itself
        for j in range(min(top_k, len(row))):
            if query_db_index is not None and row[j] == query_db_index[qi]:
                continue
            if int(row[j]) in pos:
                hits[j] += 1
                break
        if len(set(int(v) for v in row[:threshold]) & pos) > 0:
            one_pct += 1
    return hits, one_pct, evaluated


def pad_positives(positives, device=None):
    """list of sets -> (Nq, Pmax) int64 tensor padded with -1 (the device-side form of the reference's per-query lists)."""
    pmax = max((len(p) for p in positives), default=0)
    out = np.full((len(positives), max(pmax, 1)), -1, dtype=np.int64)
    for i, p in enumerate(positives):
        if p:
            out[i, : len(p)] = sorted(p)
    t = torch.from_numpy(out)
    return t.to(device) if device is not None else t


def recall_counts_device(ind, pos_padded, top_k, threshold):
    """``recall_counts`` (scene_dataset.py:1056-1081, query set disjoint from the database) for all queries at once on the
    tensor's device: ind (Nq,K) retrieved indices, pos_padded (Nq,Pmax) int64 with -1 padding.
    Returns an int64 tensor [hits_at_rank_0 .. hits_at_rank_{top_k-1}, one_percent_hits, evaluated]."""
    ind = ind.long()
    has_pos = (pos_padded >= 0).any(1)                                              # :1044-1045 queries without positives are skipped
    kk = min(top_k, ind.shape[1])
    hit = (ind[:, :kk, None] == pos_padded[:, None, :]).any(2) & has_pos[:, None]     # (Nq, kk)
    any_hit = hit.any(1)
    first = torch.where(any_hit, hit.float().argmax(1), torch.full_like(any_hit, kk, dtype=torch.long))
    hits = torch.bincount(first, minlength=kk + 1)[:kk]
    if kk < top_k:
        hits = torch.cat([hits, hits.new_zeros(top_k - kk)])
    th = min(threshold, ind.shape[1])
    one_pct = ((ind[:, :th, None] == pos_padded[:, None, :]).any(2).any(1) & has_pos).sum()
    return torch.cat([hits, one_pct.view(1), has_pos.sum().view(1)])


def evaluate_recall(db_desc, query_desc, positives, top_k=25, group=None, topk_fn=None):
    """Recall@1..top_k (%) and top-1 % recall (%) of ``query_desc`` against ``db_desc`` with queries sharded over ranks.

    db_desc (Ndb,D), query_desc (Nq,D): full matrices present on every rank (after ``extract_descriptors``);
    positives: list (len Nq) of sets of database indices.  topk_fn: override of the GPU kernel (host-logic tests).
    Returns dict(recall (top_k,), one_percent_recall, evaluated).
    """
    rank, world = _dist_info(group)
    n_db = db_desc.shape[0]
    k, threshold = real_top_k(n_db, top_k)
    k = min(k, n_db)
    lo, hi = shard_range(query_desc.shape[0], rank, world)
    if hi > lo:
        _, ind = (topk_fn or retrieval_topk)(db_desc, query_desc[lo:hi], k)
        # first-hit counters for the whole query shard in a handful of tensor ops on the descriptors' device (the reference walks
        # the queries one by one on the host, scene_dataset.py:1040-1081)
        pos = positives[lo:hi] if torch.is_tensor(positives) else pad_positives(positives[lo:hi])
        counters = recall_counts_device(ind, pos.to(ind.device), top_k, threshold).to(db_desc.device)
    else:
        counters = torch.zeros(top_k + 2, dtype=torch.int64, device=db_desc.device)
    if world > 1:
        dist.all_reduce(counters, group=group)
    counters = counters.cpu().numpy()
    evaluated = int(counters[-1])
    recall = np.cumsum(counters[:top_k]) / float(max(evaluated, 1)) * 100.0           # :1095
    return dict(recall=recall, one_percent_recall=counters[-2] / float(max(evaluated, 1)) * 100.0, evaluated=evaluated,
                k=k, threshold=threshold)


def hard_negatives(query_desc, ref_desc, negative_indices, num_hard_neg=10):
    """GPU replacement of ``SceneDataSet.__get_hard_negatives`` (scene_dataset.py:1101-1113) for a batch of queries.

    query_desc (Q,D), ref_desc (N,D) float32 CUDA; negative_indices: per query, the reference indices that are negatives.
    Returns, per query, the ``num_hard_neg`` negatives nearest in descriptor space (ascending distance, as the KDTree query
    of the reference returns them) or ``[]`` when the query has fewer than ``num_hard_neg`` negatives.
    The negative sets are ragged: they become one (Q, N/32) bit mask and ONE launch of the masked top-k kernel ranks every
    query against its own set (the reference builds a KDTree per query).  Ties go to the lower reference index.
    """
    L.require_cuda(query_desc, ref_desc)
    Q, N = query_desc.shape[0], ref_desc.shape[0]
    if Q == 0:
        return []
    words = (N + 31) // 32
    bits = np.zeros((Q, words * 32), dtype=bool)
    enough = np.zeros(Q, dtype=bool)
    for qi, neg in enumerate(negative_indices):
        neg = np.asarray(neg, dtype=np.int64)
        enough[qi] = len(np.unique(neg)) >= num_hard_neg and len(neg) >= num_hard_neg
        if len(neg):
            bits[qi, neg] = True
    mask = np.packbits(bits.reshape(Q, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(Q, words)
    mask_t = torch.from_numpy(mask.view(np.int32).copy()).to(ref_desc.device)
    k = min(num_hard_neg, N)
    out_d = torch.empty(Q, k, dtype=torch.float32, device=ref_desc.device)
    out_i = torch.empty(Q, k, dtype=torch.int32, device=ref_desc.device)
    L.check(L.lib().pab_retrieval_topk_masked(L.ptr(ref_desc.contiguous().float()), N, L.ptr(query_desc.contiguous().float()), Q,
                                              ref_desc.shape[1], k, L.ptr(mask_t), L.ptr(out_d), L.ptr(out_i), L.stream_ptr()),
            "retrieval_topk_masked")
    ind = out_i.cpu().numpy()
    return [ind[qi].tolist() if enough[qi] else [] for qi in range(Q)]


def top_k_in_feature_space(desc, positions, r_pos, r_neg, top_k=300, k_search=1000):
    """Training branch of ``SceneDataSet.find_top_k_feat`` (scene_dataset.py:884-921) with the per-record KDTree queries
    replaced by ONE batched GPU top-k over the whole record set.

    desc (N,D) float32 CUDA global descriptors; positions (N,2) numpy northing/easting (``get_dist`` = Euclidean distance
    between them, :185-189).  For every record i its ``min(k_search, N)`` nearest records in descriptor space are walked
    in order: self skipped, geometric distance < r_pos -> positive (state 1), > r_neg -> negative (state 0), otherwise
    unknown (not used); at most ``top_k // 2`` of each class are kept; when ``top_k`` entries are collected the walk stops,
    and an entry that filled up with a single class is dropped (:913-916).  Returns (top_k_dict, stats) where
    stats = dict(n_q, n_p, n_n, n_u, n_valid) are the counters the reference prints.
    """
    L.require_cuda(desc)
    n = desc.shape[0]
    k = min(k_search, n)
    _, ind = retrieval_topk(desc, desc, k)
    ind = ind.cpu().numpy()
    positions = np.asarray(positions, dtype=np.float64)
    top_k_dict = {}
    n_q = n_p = n_n = n_u = n_valid = 0
    half = top_k // 2
    for i in range(n):
        cur_p = cur_n = 0
        entry = {"top_k": [], "state": []}
        top_k_dict[i] = entry
        has_pos = False
        d = np.linalg.norm(positions[ind[i]] - positions[i], axis=1)
        for j, dij in zip(ind[i], d):
            j = int(j)
            if j == i:
                continue
            if dij < r_pos:
                if cur_p == half:
                    continue
                entry["top_k"].append(j); entry["state"].append(1)
                n_p += 1; cur_p += 1
                has_pos = True
            elif dij > r_neg:
                if cur_n == half:
                    continue
                entry["top_k"].append(j); entry["state"].append(0)
                n_n += 1; cur_n += 1
            else:
                n_u += 1
            if cur_p + cur_n == top_k:
                if cur_p == 0 or cur_n == 0:
                    del top_k_dict[i]
                break
        n_q += 1
        if has_pos:
            n_valid += 1
    return top_k_dict, dict(n_q=n_q, n_p=n_p, n_n=n_n, n_u=n_u, n_valid=n_valid)
