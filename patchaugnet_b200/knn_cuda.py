"""KNN_CUDA drop-in — mirror of ``libs/KNN_CUDA/knn_cuda/__init__.py`` without the import-time JIT build
(knn_cuda/__init__.py:10-38 compiles and asserts CUDA on import).

``knn(ref, query, k) -> (dist (k,nq), ind (k,nq) int64 0-based)`` for ref (dim,nr) / query (dim,nq), and
``KNN(k, transpose_mode).forward(ref, query) -> (D, I)`` with the reference's batching and transposition
(knn_cuda/__init__.py:41-74).  ``_knn.knn`` keeps the raw 1-based pybind contract (knn.cpp:23-56).
"""
import torch
import torch.nn as nn

from . import _lib as L

__version__ = "0.2"


class _knn:
    """Stand-in for the JIT-built extension object: _knn.knn(ref, query, k) -> [dist (k,nq), ind (k,nq) 1-based]."""

    @staticmethod
    def knn(ref, query, k):
        L.require_cuda(ref, query)
        if ref.dtype != torch.float32 or query.dtype != torch.float32:
            raise TypeError("ref/query must be float32")            # CHECK_TYPE, knn.cpp:7
        if not (ref.is_contiguous() and query.is_contiguous()):
            raise ValueError("ref/query must be contiguous")        # CHECK_CONTIGUOUS, knn.cpp:6
        dim, nr = ref.shape
        nq = query.shape[1]
        dist = torch.empty(k, nq, dtype=torch.float32, device=ref.device)
        ind = torch.empty(k, nq, dtype=torch.int64, device=ref.device)
        L.check(L.lib().pab_knn(L.ptr(ref), nr, L.ptr(query), nq, dim, k, L.ptr(dist), L.ptr(ind), L.stream_ptr()), "knn")
        return [dist, ind]


def knn(ref, query, k):
    d, i = _knn.knn(ref, query, k)
    i -= 1
    return d, i


def _T(t, mode=False):
    return t.transpose(0, 1).contiguous() if mode else t


class KNN(nn.Module):
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            D, I = [], []
            for bi in range(ref.size(0)):
                r, q = _T(ref[bi], self._t), _T(query[bi], self._t)
                d, i = knn(r.float().contiguous(), q.float().contiguous(), self.k)
                D.append(_T(d, self._t))
                I.append(_T(i, self._t))
            return torch.stack(D, dim=0), torch.stack(I, dim=0)
