"""Overlap indices and patch-feature-contrast (a2b) triplets — the step after the hot path in training.

Reference:
  * file format ``{dataset_type}_overlap_indices_{query}.pb`` = message ``QueryOverlapIndices`` of
    ``datasets/query_pos_neg_dataset.proto:14-30``, read by ``SceneDataSet.get_overlap_indices``
    (``datasets/scene_dataset.py:278-297``);
  * the a2b selection loop ``place_recognition/train_place_recognition.py:308-385`` — per (query, positive) pair a numpy
    ``where`` / ``isin`` per overlap entry and one ``index_select`` + H2D copy per selected triplet.

Here the ``.pb`` bytes are decoded straight into CSR arrays (no protobuf runtime, no per-entry Python objects), the
selection runs as ONE kernel launch for all pairs of the step (``pab_patch_triplets``, csrc/patch_pairs.cu), and the
contrastive loss is evaluated on batched gathers — no host round trip except the per-pair triplet counts.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L


# ---- proto3 wire format ------------------------------------------------------------------------------------------
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf, pos, end):
    """Yield (field number, wire type, value) of one message; value = int (varint / fixed) or (start, stop) (length-delimited)."""
    while pos < end:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = (pos, pos + n)
            pos += n
        elif wt == 5:
            v = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        elif wt == 1:
            v = int.from_bytes(buf[pos:pos + 8], "little")
            pos += 8
        else:
            raise ValueError(f"unsupported wire type {wt}")
        if pos > end:
            raise ValueError("truncated message")
        yield field, wt, v


def _repeated_uint32(buf, wt, v, out):
    if wt == 0:                                   # unpacked element
        out.append(v)
    else:                                         # packed (proto3 default)
        pos, end = v
        while pos < end:
            x, pos = _varint(buf, pos)
            out.append(x)


class OverlapEntries:
    """The ``repeated Uint32Pair`` of one (query, positive) pair as CSR arrays (int32)."""

    def __init__(self, idx1, near_ptr, near, far_ptr, far, bad_ptr, bad):
        self.idx1 = np.asarray(idx1, dtype=np.int32)
        self.near_ptr, self.near = np.asarray(near_ptr, dtype=np.int32), np.asarray(near, dtype=np.int32)
        self.far_ptr, self.far = np.asarray(far_ptr, dtype=np.int32), np.asarray(far, dtype=np.int32)
        self.bad_ptr, self.bad = np.asarray(bad_ptr, dtype=np.int32), np.asarray(bad, dtype=np.int32)

    def __len__(self):
        return len(self.idx1)

    @classmethod
    def from_lists(cls, entries):
        """entries: iterable of (idx1, near_indices2, far_indices2, bad_far_indices2)."""
        idx1, near, far, bad = [], [], [], []
        near_ptr, far_ptr, bad_ptr = [0], [0], [0]
        for i1, ne, fa, ba in entries:
            idx1.append(i1)
            near += list(ne); far += list(fa); bad += list(ba)
            near_ptr.append(len(near)); far_ptr.append(len(far)); bad_ptr.append(len(bad))
        return cls(idx1, near_ptr, near, far_ptr, far, bad_ptr, bad)

    def far_lists(self, hard_only):
        """The list ``np.isin`` is given for the negatives (train_place_recognition.py:347-360), as CSR.

        hard_only (``epoch > hard_neg_epoch_for_patch_align and use_hard_negative_patch_mining``): bad_far_indices2.
        Otherwise the reference loops ``for far_i in range(0, len(t), 2): list_far_indices = t[far_i]`` over
        t = far_indices2 + bad_far_indices2, which leaves the LAST EVEN-POSITION ELEMENT (a scalar) — reproduced as is.
        """
        if hard_only:
            return self.bad_ptr, self.bad
        ptr, val = [0], []
        for e in range(len(self)):
            t = np.concatenate([self.far[self.far_ptr[e]:self.far_ptr[e + 1]], self.bad[self.bad_ptr[e]:self.bad_ptr[e + 1]]])
            if len(t):
                val.append(int(t[2 * ((len(t) - 1) // 2)]))
            ptr.append(len(val))
        return np.asarray(ptr, dtype=np.int32), np.asarray(val, dtype=np.int32)


def _parse_pair(buf, span):
    idx1, near, far, bad = 0, [], [], []
    for field, wt, v in _fields(buf, *span):
        if field == 1:
            idx1 = v
        elif field == 2:
            _repeated_uint32(buf, wt, v, near)
        elif field == 3:
            _repeated_uint32(buf, wt, v, far)
        elif field == 4:
            _repeated_uint32(buf, wt, v, bad)
    return idx1, near, far, bad


def parse_query_overlap_indices(data):
    """bytes of a ``QueryOverlapIndices`` message -> (query_idx, {positive_idx: OverlapEntries}) (``overlap_indices`` only,
    like the reader at scene_dataset.py:290-291; ``inv_overlap_indices`` is skipped)."""
    buf = memoryview(bytes(data))
    query_idx, out = 0, {}
    for field, wt, v in _fields(buf, 0, len(buf)):
        if field == 1 and wt == 0:
            query_idx = v
        elif field == 2 and wt == 2:
            positive_idx, entries = 0, []
            for f2, wt2, v2 in _fields(buf, *v):
                if f2 == 2 and wt2 == 0:
                    positive_idx = v2
                elif f2 == 3 and wt2 == 2:
                    entries.append(_parse_pair(buf, v2))
            out[positive_idx] = OverlapEntries.from_lists(entries)
    return query_idx, out


def get_overlap_indices(data, query_idx_in_dataset, positive_indices):
    """``SceneDataSet.get_overlap_indices`` (scene_dataset.py:278-297) on the file's bytes: {(0, i+1): entries of positive i}."""
    _, per_pos = parse_query_overlap_indices(data)
    return {(0, i + 1): per_pos[p] for i, p in enumerate(positive_indices)}


def _put_varint(out, x):
    while x >= 0x80:
        out.append((x & 0x7F) | 0x80)
        x >>= 7
    out.append(x)


def encode_query_overlap_indices(query_idx, per_positive):
    """Inverse of the parser (packed repeated fields, proto3): {positive_idx: [(idx1, near, far, bad), ...]} -> bytes.
    Used to write synthetic ``.pb`` files; checked against the protobuf runtime in tests/test_host_cpu.py."""
    def packed(field, vals, out):
        if len(vals):
            body = bytearray()
            for x in vals:
                _put_varint(body, int(x))
            _put_varint(out, (field << 3) | 2); _put_varint(out, len(body)); out.extend(body)

    msg = bytearray()
    if query_idx:
        _put_varint(msg, (1 << 3) | 0); _put_varint(msg, query_idx)
    for pos_idx, entries in per_positive.items():
        qp = bytearray()
        if pos_idx:
            _put_varint(qp, (2 << 3) | 0); _put_varint(qp, pos_idx)
        for i1, ne, fa, ba in entries:
            pr = bytearray()
            if i1:
                _put_varint(pr, (1 << 3) | 0); _put_varint(pr, int(i1))
            packed(2, ne, pr); packed(3, fa, pr); packed(4, ba, pr)
            _put_varint(qp, (3 << 3) | 2); _put_varint(qp, len(pr)); qp.extend(pr)
        _put_varint(msg, (2 << 3) | 2); _put_varint(msg, len(qp)); msg.extend(qp)
    return bytes(msg)


# ---- device-side triplet selection ---------------------------------------------------------------------------------
MAX_ENTRIES_PER_PAIR = 500        # train_place_recognition.py:331-332


def sample_entries(n_entries, rng):
    """k_list of the reference (:330-332): every entry in order, or a random sample of 500 (``random.sample``; here a numpy
    Generator — the draw itself cannot be bit-matched to Python's ``random`` state of a training run)."""
    if n_entries <= MAX_ENTRIES_PER_PAIR:
        return np.arange(n_entries, dtype=np.int32)
    return rng.choice(n_entries, MAX_ENTRIES_PER_PAIR, replace=False).astype(np.int32)


class TripletBatch:
    """All overlap entries of one training step flattened for the kernel (host arrays + device copies)."""

    def __init__(self, nn_dict, cloud_rows, hard_only=False, rng=None, device="cuda"):
        """nn_dict: {(m, n): OverlapEntries}; cloud_rows: {cloud id: row of that cloud in the (n_clouds, M) centre table}."""
        rng = rng or np.random.default_rng(0)
        self.pairs = list(nn_dict)
        pair_m, pair_n, entry_ptr, idx1 = [], [], [0], []
        near_ptr, near, far_ptr, far = [0], [], [0], []
        for (m, n) in self.pairs:
            ent = nn_dict[m, n]
            fptr, fval = ent.far_lists(hard_only)
            for e in sample_entries(len(ent), rng):
                idx1.append(int(ent.idx1[e]))
                near += ent.near[ent.near_ptr[e]:ent.near_ptr[e + 1]].tolist()
                far += fval[fptr[e]:fptr[e + 1]].tolist()
                near_ptr.append(len(near)); far_ptr.append(len(far))
            pair_m.append(cloud_rows[m]); pair_n.append(cloud_rows[n]); entry_ptr.append(len(idx1))
        as_i32 = lambda x: np.asarray(x, dtype=np.int32)
        self.host = dict(pair_m=as_i32(pair_m), pair_n=as_i32(pair_n), entry_ptr=as_i32(entry_ptr), idx1=as_i32(idx1),
                         near_ptr=as_i32(near_ptr), near=as_i32(near), far_ptr=as_i32(far_ptr), far=as_i32(far))
        # ONE pinned staging buffer and one H2D copy for the whole step (the reference: thousands of 1-element copies)
        sizes = [v.size for v in self.host.values()]
        flat = torch.empty(max(1, sum(sizes)), dtype=torch.int32).pin_memory() if torch.cuda.is_available() else \
            torch.empty(max(1, sum(sizes)), dtype=torch.int32)
        off = 0
        for v in self.host.values():
            flat[off:off + v.size] = torch.from_numpy(v)
            off += v.size
        dflat = flat.to(device, non_blocking=True)
        self.dev, off = {}, 0
        for (k, v) in self.host.items():
            self.dev[k] = dflat[off:off + v.size]
            off += v.size
        self.max_entries = int(np.diff(self.host["entry_ptr"]).max()) if self.pairs else 0


def select_triplets(batch, centers, seed=0, max_out=None):
    """centers: (n_clouds, M) int32 CUDA — level-0 centre indices (``center_indices`` of the forward's patch_recon dict).
    Returns (idx1, pos, neg) (n_pairs, max_out) int32 and counts (n_pairs,) int32, all on the device."""
    L.require_cuda(centers)
    centers = centers.contiguous()
    assert centers.dtype == torch.int32 and centers.dim() == 2
    n_pairs, M = len(batch.pairs), centers.shape[1]
    if max_out is None:                                  # positives per entry <= centres; usually a handful
        max_out = max(1, min(batch.max_entries * M, 4 * int(batch.host["near"].size // max(1, n_pairs)) + 64))
    d, p = batch.dev, L.ptr
    while True:
        out = torch.empty(3, max(1, n_pairs), max_out, dtype=torch.int32, device=centers.device)
        count = torch.zeros(max(1, n_pairs), dtype=torch.int32, device=centers.device)
        L.check(L.lib().pab_patch_triplets(n_pairs, M, p(centers), p(d["pair_m"]), p(d["pair_n"]), p(d["entry_ptr"]), batch.max_entries,
                                           p(d["idx1"]), p(d["near_ptr"]), p(d["near"]), p(d["far_ptr"]), p(d["far"]),
                                           C.c_ulonglong(seed), max_out, p(out[0]), p(out[1]), p(out[2]), p(count), L.stream_ptr()),
                "pab_patch_triplets")
        need = int(count.max().item()) if n_pairs else 0     # the one host read-back of the step
        if need <= max_out:
            return out[0][:n_pairs], out[1][:n_pairs], out[2][:n_pairs], count[:n_pairs]
        max_out = need


def patch_feature_contrast_loss(nn_dict, cloud_indices, center_indices, patch_features, margin, hard_only=False, seed=0, rng=None):
    """``cur_loss['patch_recon_a2b']`` of train_place_recognition.py:308-385.

    nn_dict {(m, n): OverlapEntries}; cloud_indices / center_indices / patch_features as in the forward's patch_recon dict
    (lists per related cloud: cloud id, (1, M) centre indices, (M, D) patch features).  Returns (loss, n_pairs_used);
    loss = mean over pairs that produced triplets of contrastive_loss(q, p, n, margin) (pointnetvlad_loss.py:170-186).
    """
    rows = {int(c): k for k, c in reversed(list(enumerate(cloud_indices)))}      # first k with cloud_indices[k] == m (:314-318)
    centers = torch.stack([c.reshape(-1) for c in center_indices]).to(torch.int32)
    feats = torch.stack(list(patch_features))                                    # (n_clouds, M, D)
    batch = TripletBatch(nn_dict, rows, hard_only=hard_only, rng=rng, device=centers.device)
    i1, ip, ineg, count = select_triplets(batch, centers, seed=seed)
    n_pairs, M, D = len(batch.pairs), centers.shape[1], feats.shape[2]
    if n_pairs == 0:
        return feats.new_zeros(()), 0
    T = i1.shape[1]
    valid = torch.arange(T, device=count.device)[None, :] < count[:, None]                      # (P, T)
    rm, rn = batch.dev["pair_m"].long()[:, None] * M, batch.dev["pair_n"].long()[:, None] * M
    flat = feats.reshape(-1, D)
    sel = valid.reshape(-1).nonzero().squeeze(1)
    q = flat[(rm + i1.long()).reshape(-1)[sel]]
    pz = flat[(rn + ip.long()).reshape(-1)[sel]]
    ng = flat[(rn + ineg.long()).reshape(-1)[sel]]
    seg = (torch.arange(n_pairs, device=count.device)[:, None].expand(n_pairs, T)).reshape(-1)[sel]
    qp = F.pairwise_distance(q, pz).pow(2)
    qn = torch.clamp(margin - F.pairwise_distance(q, ng), min=0.0).pow(2)
    per_pair = torch.zeros(n_pairs, device=flat.device, dtype=flat.dtype).index_add_(0, seg, qp + qn)
    cnt = count.to(flat.dtype)
    used = count > 0
    loss = (per_pair[used] / cnt[used]).sum()
    n_used = int(used.sum().item())
    return (loss / n_used if n_used else loss), n_used
