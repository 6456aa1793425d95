#!/usr/bin/env python
"""Retrieval top-k timing: single-pass vs database-split kernel, 250 and 2000 queries x 10k database, k = 101."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patchaugnet_b200 import _lib as L, retrieval
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
db = torch.nn.functional.normalize(torch.randn(10000, 256, device="cuda"))
out = {}
for nq in (250, 2000):
    q = torch.nn.functional.normalize(torch.randn(nq, 256, device="cuda"))
    d = torch.empty(nq, 101, device="cuda"); i = torch.empty(nq, 101, dtype=torch.int32, device="cuda")
    out[f"single_nq{nq}"] = round(timed(lambda: L.lib().pab_retrieval_topk(L.ptr(db), 10000, L.ptr(q), nq, 256, 101, L.ptr(d), L.ptr(i), L.stream_ptr())), 3)
    out[f"split_nq{nq}"] = round(timed(lambda: retrieval.retrieval_topk(db, q, 101)), 3)
print(json.dumps(out))
