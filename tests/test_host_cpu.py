"""CPU tier: host-side logic — BN folding, sharding, recall bookkeeping, and the 2-rank (gloo) shard + all-gather +
recall path, which must reproduce the 1-rank result exactly."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from patchaugnet_b200 import retrieval
from patchaugnet_b200.engine import _Layers, _fold_bn


def test_folded_shared_mlp_equals_eval_forward():
    net = util.build_network()
    mlp = net.backbone.SA_modules[1].mlps[0]                 # [67, 64, 64, 256]
    layers = _Layers(mlp, "cpu")
    x = torch.randn(50, 67)
    y = x
    for i in range(layers.n):
        wt, sh = layers.tensors[2 * i], layers.tensors[2 * i + 1]
        assert wt.shape[0] % 4 == 0 and (wt[layers.spec[i][0]:] == 0).all()
        y = torch.relu(torch.nn.functional.pad(y, (0, wt.shape[0] - y.shape[1])) @ wt + sh)
    ref = mlp(x.t()[None, :, :, None]).squeeze(-1)[0].t()
    assert torch.allclose(y, ref, atol=1e-5, rtol=1e-5)
    # folding must not touch the module's own parameters (state_dict is API)
    assert "layer0.conv.weight" in mlp.state_dict() and mlp.layer0.conv.weight.shape == (64, 67, 1, 1)


def test_fold_bn_matches_batchnorm_eval():
    bn = torch.nn.BatchNorm1d(8).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.data.normal_(); bn.bias.data.normal_()
    x = torch.randn(5, 8)
    s, t = _fold_bn(bn)
    assert torch.allclose(x * s + t, bn(x), atol=1e-6)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10000, 10001):
        for w in (1, 2, 4, 8):
            spans = [retrieval.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_real_top_k_rule():
    assert retrieval.real_top_k(10000, 25) == (101, 100)      # max(26, round(N/100)+1), scene_dataset.py:1026-1029
    assert retrieval.real_top_k(300, 25) == (26, 3)
    assert retrieval.real_top_k(10, 25) == (26, 1)


def test_recall_counts_first_hit_rule():
    ind = np.array([[5, 3, 9, 1], [2, 2, 2, 2], [7, 8, 9, 0]])
    pos = [{9, 1}, set(), {0}]
    hits, one_pct, ev = retrieval.recall_counts(ind, pos, top_k=3, threshold=2)
    assert ev == 2 and hits.tolist() == [0, 0, 1] and one_pct == 0            # query 2's positive sits at rank 4 > top_k
    hits, one_pct, ev = retrieval.recall_counts(ind, pos, top_k=4, threshold=3)
    assert hits.tolist() == [0, 0, 1, 1] and one_pct == 1


def _cpu_topk(db, q, k):
    d = torch.cdist(q.double(), db.double())
    order = torch.argsort(d, dim=1, stable=True)[:, :k]
    return torch.gather(d, 1, order).float(), order.int()


def _fake_extract(x):                      # deterministic stand-in for the network: (b,1,N,3) -> (b,8)
    x = x.squeeze(1)
    feats = torch.cat([x.mean(1), x.std(1), x.abs().max(1)[0][:, :2]], 1)
    return torch.nn.functional.normalize(feats)


def _world(rank, world, port, clouds, queries, positives, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = retrieval.extract_descriptors(_fake_extract, clouds, batch_size=3, dim=8)
        qd = retrieval.extract_descriptors(_fake_extract, queries, batch_size=3, dim=8)
        res = retrieval.evaluate_recall(db, qd, positives, top_k=5, topk_fn=_cpu_topk)
        if rank == 0:
            out.put((db.numpy(), res["recall"], res["one_percent_recall"], res["evaluated"]))
    finally:
        dist.destroy_process_group()


def test_device_recall_counters_equal_the_per_query_loop():
    """retrieval.recall_counts_device (tensor ops, all queries at once) against the restated per-query loop of
    scene_dataset.py:1056-1081 on random rankings: empty positive sets, several positives, hits beyond top_k."""
    rng = np.random.default_rng(11)
    nq, ndb, K, top_k, thr = 300, 500, 40, 25, 5
    ind = np.stack([rng.permutation(ndb)[:K] for _ in range(nq)]).astype(np.int32)
    positives = []
    for q in range(nq):
        r = rng.random()
        positives.append(set() if r < 0.1 else set(rng.choice(ndb, size=int(rng.integers(1, 6)), replace=False).tolist()))
    hits, one_pct, ev = retrieval.recall_counts(ind, positives, top_k, thr)
    got = retrieval.recall_counts_device(torch.from_numpy(ind), retrieval.pad_positives(positives), top_k, thr).numpy()
    assert np.array_equal(got[:top_k], hits) and got[top_k] == one_pct and got[top_k + 1] == ev
    # fewer retrieved neighbours than top_k
    hits2, one2, ev2 = retrieval.recall_counts(ind[:, :10], positives, top_k, thr)
    got2 = retrieval.recall_counts_device(torch.from_numpy(ind[:, :10]), retrieval.pad_positives(positives), top_k, thr).numpy()
    assert np.array_equal(got2[:top_k], hits2) and got2[top_k] == one2 and got2[top_k + 1] == ev2


def test_two_rank_gloo_matches_single_rank():
    g = torch.Generator().manual_seed(0)
    clouds = torch.rand(11, 64, 3, generator=g)                       # 11 places: uneven shards on 2 ranks
    queries = clouds[:7] + 0.01 * torch.randn(7, 64, 3, generator=g)
    positives = [{i} for i in range(7)]
    db1 = retrieval.extract_descriptors(_fake_extract, clouds, batch_size=3, dim=8)
    q1 = retrieval.extract_descriptors(_fake_extract, queries, batch_size=3, dim=8)
    res1 = retrieval.evaluate_recall(db1, q1, positives, top_k=5, topk_fn=_cpu_topk)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_world, args=(r, 2, port, clouds, queries, positives, out)) for r in range(2)]
    for p in procs:
        p.start()
    db2, recall2, one2, ev2 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(db2, db1.numpy())                           # gathered database identical to the 1-rank run
    assert np.array_equal(recall2, res1["recall"]) and one2 == res1["one_percent_recall"] and ev2 == res1["evaluated"] == 7


# ---- overlap indices (.pb) — SURVEY 8(f) rank 3 --------------------------------------------------------------------
def _overlap_proto_classes():
    """QueryOverlapIndices & co. of datasets/query_pos_neg_dataset.proto:14-30, declared through the protobuf runtime's
    descriptor API (no protoc in the image) — the independent encoder/decoder the hand-written wire parser is pinned to."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="qpn_test.proto", package="p2m.base_type", syntax="proto3")
    U32, MSG = descriptor_pb2.FieldDescriptorProto.TYPE_UINT32, descriptor_pb2.FieldDescriptorProto.TYPE_MESSAGE
    OPT, REP = descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL, descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED
    m = fd.message_type.add(name="Uint32Pair")
    m.field.add(name="idx1", number=1, type=U32, label=OPT)
    for i, n in enumerate(["near_indices2", "far_indices2", "bad_far_indices2"]):
        m.field.add(name=n, number=2 + i, type=U32, label=REP)
    m = fd.message_type.add(name="QueryPosOverlapIndices")
    m.field.add(name="positive_idx", number=2, type=U32, label=OPT)
    m.field.add(name="overlap_indices", number=3, type=MSG, label=REP, type_name=".p2m.base_type.Uint32Pair")
    m.field.add(name="inv_overlap_indices", number=4, type=MSG, label=REP, type_name=".p2m.base_type.Uint32Pair")
    m = fd.message_type.add(name="QueryOverlapIndices")
    m.field.add(name="query_idx", number=1, type=U32, label=OPT)
    m.field.add(name="qp_overlap_indices", number=2, type=MSG, label=REP, type_name=".p2m.base_type.QueryPosOverlapIndices")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("p2m.base_type.QueryOverlapIndices"))


def _random_overlap(rng, n_pos=3, n_entries=40, n_points=4096):
    per_pos = {}
    for p in rng.choice(5000, n_pos, replace=False):
        ents = []
        for _ in range(int(rng.integers(0, n_entries))):
            ents.append((int(rng.integers(0, n_points)), rng.integers(0, n_points, rng.integers(0, 12)).tolist(),
                         rng.integers(0, n_points, rng.integers(0, 6)).tolist(), rng.integers(0, n_points, rng.integers(0, 4)).tolist()))
        per_pos[int(p)] = ents
    return per_pos


def test_overlap_indices_wire_parser_matches_the_protobuf_runtime():
    from patchaugnet_b200 import overlap_indices as oi
    Q = _overlap_proto_classes()
    rng = np.random.default_rng(5)
    for trial in range(4):
        per_pos = _random_overlap(rng)
        if trial == 3:
            per_pos = {}                                        # empty file body
        msg = Q(query_idx=17 + trial)
        for p, ents in per_pos.items():
            qp = msg.qp_overlap_indices.add(positive_idx=p)
            for i1, ne, fa, ba in ents:
                qp.overlap_indices.add(idx1=i1, near_indices2=ne, far_indices2=fa, bad_far_indices2=ba)
            qp.inv_overlap_indices.add(idx1=3, near_indices2=[1, 2])         # present in real files, skipped by the reader
        data = msg.SerializeToString()
        qidx, got = oi.parse_query_overlap_indices(data)
        assert qidx == 17 + trial and set(got) == set(per_pos)
        for p, ents in per_pos.items():
            g = got[p]
            assert g.idx1.tolist() == [e[0] for e in ents]
            for k, (i1, ne, fa, ba) in enumerate(ents):
                assert g.near[g.near_ptr[k]:g.near_ptr[k + 1]].tolist() == ne
                assert g.far[g.far_ptr[k]:g.far_ptr[k + 1]].tolist() == fa
                assert g.bad[g.bad_ptr[k]:g.bad_ptr[k + 1]].tolist() == ba
        # the encoder writes what the runtime parses back to the same message (minus the inverse lists)
        back = Q()
        back.ParseFromString(oi.encode_query_overlap_indices(17 + trial, per_pos))
        assert back.query_idx == 17 + trial
        assert [q.positive_idx for q in back.qp_overlap_indices] == list(per_pos)
        for q, ents in zip(back.qp_overlap_indices, per_pos.values()):
            assert [(e.idx1, list(e.near_indices2), list(e.far_indices2), list(e.bad_far_indices2)) for e in q.overlap_indices] == ents
        # get_overlap_indices keys (scene_dataset.py:293-296)
        pos_list = list(per_pos)[::-1]
        d = oi.get_overlap_indices(data, 17 + trial, pos_list)
        assert list(d) == [(0, i + 1) for i in range(len(pos_list))]


def test_far_list_quirk_and_entry_sampling():
    from patchaugnet_b200 import overlap_indices as oi
    ents = oi.OverlapEntries.from_lists([(1, [5], [10, 11, 12], [13]), (2, [6], [], []), (3, [7], [20], []), (4, [], [30, 31], [32, 33, 34])])
    ptr, val = ents.far_lists(hard_only=False)
    # reference :352-359: t = far + bad; for far_i in range(0, len(t), 2): list_far_indices = t[far_i]  -> last even-position element
    want = []
    for far, bad in [([10, 11, 12], [13]), ([], []), ([20], []), ([30, 31], [32, 33, 34])]:
        t, lst = far + bad, []
        for far_i in range(0, len(t), 2):
            lst = t[far_i]
        want.append([lst] if t else [])
    assert [val[ptr[e]:ptr[e + 1]].tolist() for e in range(4)] == want
    ptr, val = ents.far_lists(hard_only=True)
    assert [val[ptr[e]:ptr[e + 1]].tolist() for e in range(4)] == [[13], [], [], [32, 33, 34]]
    rng = np.random.default_rng(0)
    assert oi.sample_entries(7, rng).tolist() == list(range(7))
    s = oi.sample_entries(1200, rng)
    assert len(s) == 500 and len(set(s.tolist())) == 500 and s.max() < 1200


def test_triplet_batch_flattening_matches_the_per_pair_entry_lists():
    """overlap_indices.TripletBatch (host side of pab_patch_triplets): CSR arrays of a whole step == the per-pair lists the
    restated reference loop (oracle/patch_pairs.py) consumes, in processing order, with the far-list rule applied."""
    from oracle import patch_pairs
    from patchaugnet_b200 import overlap_indices as oi
    rng = np.random.default_rng(21)
    nn_dict = {}
    for pair, ne in zip([(0, 1), (0, 2), (3, 1)], (5, 0, 620)):
        ents = [(int(rng.integers(0, 4096)), rng.integers(0, 4096, rng.integers(0, 9)).tolist(),
                 rng.integers(0, 4096, rng.integers(0, 5)).tolist(), rng.integers(0, 4096, rng.integers(0, 3)).tolist()) for _ in range(ne)]
        nn_dict[pair] = oi.OverlapEntries.from_lists(ents)
    for hard in (False, True):
        batch = oi.TripletBatch(nn_dict, {c: 10 + c for c in range(4)}, hard_only=hard, rng=np.random.default_rng(8), device="cpu")
        h = batch.host
        assert h["pair_m"].tolist() == [10, 10, 13] and h["pair_n"].tolist() == [11, 12, 11]
        assert np.diff(h["entry_ptr"]).tolist() == [5, 0, 500] and batch.max_entries == 500       # 620 entries sampled down to 500
        order_rng = np.random.default_rng(8)
        e = 0
        for (m, n) in batch.pairs:
            ent = nn_dict[m, n]
            fptr, fval = ent.far_lists(hard)
            for k in oi.sample_entries(len(ent), order_rng):
                assert h["idx1"][e] == ent.idx1[k]
                assert h["near"][h["near_ptr"][e]:h["near_ptr"][e + 1]].tolist() == ent.near[ent.near_ptr[k]:ent.near_ptr[k + 1]].tolist()
                assert h["far"][h["far_ptr"][e]:h["far_ptr"][e + 1]].tolist() == fval[fptr[k]:fptr[k + 1]].tolist()
                e += 1
        assert e == len(h["idx1"])
        for k, v in h.items():                                   # the device copies are views of one flat upload
            assert torch.equal(batch.dev[k], torch.from_numpy(v))
    # the restated loop and its counter-based draw are deterministic functions of (seed, pair, entry, j)
    centers = np.arange(1024, dtype=np.int32) * 3
    a = patch_pairs.select_pair(centers, centers, [(3, [3, 6, 9], [12, 15]), (5, [3], [12])], seed=7, pair=2)
    assert a == patch_pairs.select_pair(centers, centers, [(3, [3, 6, 9], [12, 15]), (5, [3], [12])], seed=7, pair=2)
    assert a[0] == [1, 1, 1] and a[1] == [1, 2, 3] and set(a[2]) <= {4, 5}


# ---- training step host logic (BASELINE.json configs[4]) -----------------------------------------------------------------
class _StubNet(torch.nn.Module):
    """Stands in for patch_aug_net.Network on the CPU (the real model has no CPU path): (A*18,1,N,3) -> (A*18, 8) descriptors."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.fc = torch.nn.Linear(6, 8)

    def forward(self, x, nn_dict=None, return_feat=True):
        x = x.squeeze(1)
        return torch.nn.functional.normalize(self.fc(torch.cat([x.mean(1), x.std(1)], 1)))


def _train_world(rank, world, port, feed, out):
    import torch.distributed as dist
    from patchaugnet_b200 import training
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _StubNet()
        model = training.build_ddp(net, torch.device("cpu"))
        assert isinstance(model, torch.nn.parallel.DistributedDataParallel)
        per = feed.shape[0] // world
        step = training.TrainStep(model, torch.optim.SGD(model.parameters(), lr=0.1), n_anchors=per // training.CLOUDS_PER_ANCHOR,
                                  use_patch_recon=False)
        step(feed[rank * per:(rank + 1) * per])
        if rank == 0:
            out.put({k: v.detach().numpy() for k, v in net.state_dict().items()})
    finally:
        dist.destroy_process_group()


def test_training_step_host_logic_and_two_rank_ddp_equals_single_rank():
    from patchaugnet_b200 import training
    d = training.make_nn_dict(2)
    assert list(d) == [(0, 1), (0, 2), (18, 19), (18, 20)]                       # scene_dataset.py:293-296 key layout
    q, pos, neg, other = training.split_descriptors(torch.arange(2 * 18 * 4.0).view(36, 4), 2)
    assert q.shape == (2, 1, 4) and pos.shape == (2, 2, 4) and neg.shape == (2, 14, 4) and other.shape == (2, 1, 4)
    assert torch.equal(other[1, 0], torch.arange(2 * 18 * 4.0).view(36, 4)[35])
    g = torch.Generator().manual_seed(1)
    feed = torch.rand(4 * 18, 1, 32, 3, generator=g)
    # single rank, 4 anchors
    net = _StubNet()
    step = training.TrainStep(net, torch.optim.SGD(net.parameters(), lr=0.1), n_anchors=4, use_patch_recon=False)
    loss, terms = step(feed)
    assert torch.isfinite(loss) and set(terms) == {"place_recognition"}
    want = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    # two ranks, 2 anchors each: DDP's gradient average of per-rank means == gradient of the global mean
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_train_world, args=(r, 2, port, feed, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in want:
        assert np.allclose(got[k], want[k], atol=1e-6), k
