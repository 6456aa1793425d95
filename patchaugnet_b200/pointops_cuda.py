"""Drop-in for the reference's pybind module ``pointops_cuda`` (libs/pointops/src/pointops_api.cpp:15-40).

Same 17 function names, same argument order, same "caller allocates, function fills" contract — but each call goes
through the C ABI of libpatchaug_b200.so (include/patchaug_b200.h) on PyTorch's CURRENT stream and raises a Python
exception on failure instead of ``exit(-1)``.  Put ``<repo>/dropin`` on ``sys.path`` to let the reference's own
``libs/pointops/functions/pointops.py`` import this module unchanged.
"""
import torch

from . import _lib as L


def _chk(t, dtype, name):
    L.require_cuda(t)
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")   # CHECK_CONTIGUOUS, cuda_utils.h:10


def _f(t, name):
    _chk(t, torch.float32, name)
    return L.ptr(t)


def _i(t, name):
    _chk(t, torch.int32, name)
    return L.ptr(t)


def _l(t, name):
    _chk(t, torch.int64, name)
    return L.ptr(t)


def ballquery_cuda(b, n, m, radius, nsample, new_xyz, xyz, idx):
    L.check(L.lib().pab_ballquery(b, n, m, radius, nsample, _f(new_xyz, "new_xyz"), _f(xyz, "xyz"), _i(idx, "idx"),
                                  L.stream_ptr()), "ballquery_cuda")


def knnquery_cuda(b, n, m, nsample, xyz, new_xyz, idx, dist2):
    L.check(L.lib().pab_knnquery(b, n, m, nsample, _f(xyz, "xyz"), _f(new_xyz, "new_xyz"), _i(idx, "idx"),
                                 _f(dist2, "dist2") if dist2 is not None else L.ptr(None), L.stream_ptr()), "knnquery_cuda")


def grouping_forward_cuda(b, c, n, m, nsample, points, idx, out):
    L.check(L.lib().pab_grouping_forward(b, c, n, m, nsample, _f(points, "points"), _i(idx, "idx"), _f(out, "out"),
                                         L.stream_ptr()), "grouping_forward_cuda")


DETERMINISTIC_BACKWARD = True      # scatter-adds as ordered gathers over an inverted index (scatter.cu); False: fp32 atomics like the reference
_MAX_L = 32768


_INVERSE_CACHE = {}      # id(idx) -> (weakref(idx), idx._version, (b, n, L), stream, workspace): the inverse index of a live index tensor


def _scatter(b, c, n, Lc, grad_out, idx, weight, grad_points, what, gdiv=1):
    """grad_points[b,ch,idx[b,e]] += grad_out[b,ch,e/gdiv] * weight[b,e], deterministic.  Returns False when the shape is out of
    range.  The inverse index is cached per index TENSOR OBJECT (weak reference + version counter): the coordinate and the
    feature grouping of an SA module share one idx, so their backward passes share one inversion."""
    import weakref
    if not DETERMINISTIC_BACKWARD or Lc > _MAX_L or Lc == 0 or b == 0 or c == 0:
        return False
    st = L.stream_ptr()
    key, build, ws = id(idx), 1, None
    hit = _INVERSE_CACHE.get(key)
    if hit is not None and hit[0]() is idx and hit[1] == idx._version and hit[2] == (b, n, Lc) and hit[3] == st.value:
        ws, build = hit[4], 0
    if ws is None:
        for k in [k for k, v in _INVERSE_CACHE.items() if v[0]() is None]:        # entries of dead tensors
            del _INVERSE_CACHE[k]
        ws = torch.empty(L.lib().pab_scatter_workspace_bytes(b, n, Lc), dtype=torch.uint8, device=grad_out.device)
        _INVERSE_CACHE[key] = (weakref.ref(idx), idx._version, (b, n, Lc), st.value, ws)
    L.check(L.lib().pab_scatter_add_deterministic_ex(b, c, n, Lc, gdiv, _f(grad_out, "grad_out"), _i(idx, "idx"),
                                                     _f(weight, "weight") if weight is not None else L.ptr(None),
                                                     _f(grad_points, "grad_points"), L.ptr(ws), build, st), what)
    return True


def grouping_backward_cuda(b, c, n, m, nsample, grad_out, idx, grad_points):
    if _scatter(b, c, n, m * nsample, grad_out, idx, None, grad_points, "grouping_backward_cuda"):
        return
    L.check(L.lib().pab_grouping_backward(b, c, n, m, nsample, _f(grad_out, "grad_out"), _i(idx, "idx"),
                                          _f(grad_points, "grad_points"), L.stream_ptr()), "grouping_backward_cuda")


def grouping_int_forward_cuda(b, c, n, m, nsample, points, idx, out):
    L.check(L.lib().pab_grouping_int_forward(b, c, n, m, nsample, _l(points, "points"), _i(idx, "idx"), _l(out, "out"),
                                             L.stream_ptr()), "grouping_int_forward_cuda")


def gathering_forward_cuda(b, c, n, m, points, idx, out):
    L.check(L.lib().pab_gathering_forward(b, c, n, m, _f(points, "points"), _i(idx, "idx"), _f(out, "out"),
                                          L.stream_ptr()), "gathering_forward_cuda")


def gathering_backward_cuda(b, c, n, m, grad_out, idx, grad_points):
    if _scatter(b, c, n, m, grad_out, idx, None, grad_points, "gathering_backward_cuda"):
        return
    L.check(L.lib().pab_gathering_backward(b, c, n, m, _f(grad_out, "grad_out"), _i(idx, "idx"),
                                           _f(grad_points, "grad_points"), L.stream_ptr()), "gathering_backward_cuda")


def furthestsampling_cuda(b, n, m, xyz, temp, idx):
    L.check(L.lib().pab_furthestsampling(b, n, m, _f(xyz, "xyz"), _f(temp, "temp") if temp is not None else L.ptr(None),
                                         _i(idx, "idx"), L.stream_ptr()), "furthestsampling_cuda")


def nearestneighbor_cuda(b, n, m, unknown, known, dist2, idx):
    L.check(L.lib().pab_nearestneighbor(b, n, m, _f(unknown, "unknown"), _f(known, "known"), _f(dist2, "dist2"),
                                        _i(idx, "idx"), L.stream_ptr()), "nearestneighbor_cuda")


def interpolation_forward_cuda(b, c, m, n, points, idx, weight, out):
    L.check(L.lib().pab_interpolation_forward(b, c, m, n, _f(points, "points"), _i(idx, "idx"), _f(weight, "weight"),
                                              _f(out, "out"), L.stream_ptr()), "interpolation_forward_cuda")


def interpolation_backward_cuda(b, c, n, m, grad_out, idx, weight, grad_points):
    # entry e = 3 j + t scatters grad_out[b, ch, j] * weight[b, j, t] to known point idx[b, j, t]
    if _scatter(b, c, m, 3 * n, grad_out, idx, weight, grad_points, "interpolation_backward_cuda", gdiv=3):
        return
    L.check(L.lib().pab_interpolation_backward(b, c, n, m, _f(grad_out, "grad_out"), _i(idx, "idx"), _f(weight, "weight"),
                                               _f(grad_points, "grad_points"), L.stream_ptr()), "interpolation_backward_cuda")


def labelstat_idx_cuda(b, n, m, nsample, nclass, label_stat, idx, new_label_stat):
    L.check(L.lib().pab_labelstat_idx(b, n, m, nsample, nclass, _i(label_stat, "label_stat"), _i(idx, "idx"),
                                      _i(new_label_stat, "new_label_stat"), L.stream_ptr()), "labelstat_idx_cuda")


def labelstat_ballrange_cuda(b, n, m, radius, nclass, new_xyz, xyz, label_stat, new_label_stat):
    L.check(L.lib().pab_labelstat_ballrange(b, n, m, radius, nclass, _f(new_xyz, "new_xyz"), _f(xyz, "xyz"),
                                            _i(label_stat, "label_stat"), _i(new_label_stat, "new_label_stat"),
                                            L.stream_ptr()), "labelstat_ballrange_cuda")


def labelstat_and_ballquery_cuda(b, n, m, radius, nsample, nclass, new_xyz, xyz, label_stat, idx, new_label_stat):
    L.check(L.lib().pab_labelstat_and_ballquery(b, n, m, radius, nsample, nclass, _f(new_xyz, "new_xyz"), _f(xyz, "xyz"),
                                                _i(label_stat, "label_stat"), _i(idx, "idx"),
                                                _i(new_label_stat, "new_label_stat"), L.stream_ptr()),
            "labelstat_and_ballquery_cuda")


def featuredistribute_cuda(b, n, m, max_xyz, xyz, distribute_idx):
    L.check(L.lib().pab_featuredistribute(b, n, m, _f(max_xyz, "max_xyz"), _f(xyz, "xyz"),
                                          _i(distribute_idx, "distribute_idx"), L.stream_ptr()), "featuredistribute_cuda")


def featuregather_forward_cuda(b, n, m, c, max_feature, distribute_idx, distribute_feature):
    L.check(L.lib().pab_featuregather_forward(b, n, m, c, _f(max_feature, "max_feature"),
                                              _i(distribute_idx, "distribute_idx"),
                                              _f(distribute_feature, "distribute_feature"), L.stream_ptr()),
            "featuregather_forward_cuda")


def featuregather_backward_cuda(b, n, m, c, grad_distribute_feature, distribute_idx, grad_max_feature):
    L.check(L.lib().pab_featuregather_backward(b, n, m, c, _f(grad_distribute_feature, "grad_distribute_feature"),
                                               _i(distribute_idx, "distribute_idx"),
                                               _f(grad_max_feature, "grad_max_feature"), L.stream_ptr()),
            "featuregather_backward_cuda")


__all__ = [
    "ballquery_cuda", "knnquery_cuda", "grouping_forward_cuda", "grouping_backward_cuda", "grouping_int_forward_cuda",
    "gathering_forward_cuda", "gathering_backward_cuda", "furthestsampling_cuda", "nearestneighbor_cuda",
    "interpolation_forward_cuda", "interpolation_backward_cuda", "labelstat_idx_cuda", "labelstat_ballrange_cuda",
    "labelstat_and_ballquery_cuda", "featuredistribute_cuda", "featuregather_forward_cuda", "featuregather_backward_cuda",
]
