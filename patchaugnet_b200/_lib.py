"""ctypes binding of libpatchaug_b200.so — the thin C-ABI doorway (include/patchaug_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpatchaug_b200.so")
PAB_EINVAL = -22

_lib = None

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_L = C.c_long
_SZ = C.c_size_t


class PabLayer(C.Structure):
    """pab_layer_t"""
    _fields_ = [("wt", _P), ("shift", _P), ("c_in", _I), ("c_in_pad", _I), ("c_out", _I), ("relu", _I),
                ("w_hi", _P), ("w_lo", _P), ("tc_k0", _I), ("tc_k", _I)]


# name -> (restype, argtypes); every symbol include/patchaug_b200.h declares
SIGNATURES = {
    "pab_version": (_I, []),
    "pab_num_launches": (_I, []),
    "pab_reset_launch_counter": (None, []),
    "pab_tune_fps_threads": (None, [_I]),
    "pab_tune_tensor_core": (None, [_I]),
    "pab_tune_tc_trace": (None, [_P]),
    "pab_tune_tc_max_ctas": (None, [_I]),
    "pab_tune_fps_clouds_per_cta": (None, [_I]),
    "pab_tune_fps_pruned": (None, [_I]),
    "pab_tune_attention_small": (None, [_I]),
    "pab_tune_pointwise_tc": (None, [_I]),
    "pab_tune_sa_narrow": (None, [_I, _I]),
    "pab_tune_sa_narrow_trace": (None, [_P]),
    "pab_tune_sa_narrow_dbg": (None, [_I]),
    "pab_tune_fps_exclusive": (None, [_I]),
    "pab_bn_train_workspace_bytes": (C.c_size_t, [_I]),
    "pab_bn_relu_train_forward": (_I, [_I, _I, C.c_long, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "pab_bn_relu_train_backward": (_I, [_I, _I, C.c_long, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pab_scatter_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "pab_scatter_add_deterministic": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pab_scatter_add_deterministic_ex": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P]),
    "pab_fps_clouds_per_sm": (_I, [_I]),
    "pab_furthestsampling": (_I, [_I, _I, _I, _P, _P, _P, _P]),
    "pab_gathering_forward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_gathering_backward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_knnquery": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pab_knn_index_bytes": (_SZ, [_I, _I]),
    "pab_knn_build_index": (_I, [_I, _I, _P, _P, _P]),
    "pab_knnquery_indexed": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pab_three_nn_weights_indexed": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pab_ballquery": (_I, [_I, _I, _I, _F, _I, _P, _P, _P, _P]),
    "pab_grouping_forward": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_grouping_backward": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_grouping_int_forward": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_nearestneighbor": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "pab_interpolation_forward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pab_interpolation_backward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pab_featuredistribute": (_I, [_I, _I, _I, _P, _P, _P, _P]),
    "pab_featuregather_forward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_featuregather_backward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_labelstat_idx": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_labelstat_ballrange": (_I, [_I, _I, _I, _F, _I, _P, _P, _P, _P, _P]),
    "pab_labelstat_and_ballquery": (_I, [_I, _I, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pab_chamfer_forward": (_I, [_I, _I, _P, _I, _P, _P, _P, _P, _P, _P]),
    "pab_chamfer_backward": (_I, [_I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pab_knn": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P]),
    "pab_retrieval_topk": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P]),
    "pab_retrieval_topk_masked": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "pab_retrieval_topk_workspace_bytes": (C.c_size_t, [_I, _I]),
    "pab_retrieval_topk_split": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "pab_emd_forward": (_I, [_I, _I] + [_P] * 14 + [_F, _I, _P]),
    "pab_emd_backward": (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "pab_gather_rows": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "pab_three_nn_weights": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "pab_sa_module_forward": (_I, [_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, C.POINTER(PabLayer), _I, _P, _P, _P]),
    "pab_fp_module_forward": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, C.POINTER(PabLayer), _I, _P, _P]),
    "pab_fp_module_forward_ordered": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _L, C.POINTER(PabLayer), _I, _P, _P]),
    "pab_knn_index_order": (_P, [_I, _P, C.POINTER(_L)]),
    "pab_pointwise_mlp_forward": (_I, [_I, _P, C.POINTER(PabLayer), _I, _P, _P]),
    "pab_sa_layer_workspace_bytes": (_SZ, [_I, _I, _I]),
    "pab_sa_layer_forward": (_I, [_I, _I, _I, _P, C.POINTER(PabLayer), C.POINTER(PabLayer), C.POINTER(PabLayer), _P, _P, _P]),
    "pab_sa_layer_forward_p": (_I, [_I, _I, _I, _P, C.POINTER(PabLayer), C.POINTER(PabLayer), C.POINTER(PabLayer), _P, _P, _I, _P]),
    "pab_netvlad_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "pab_netvlad_forward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _L, _L, _P, _P]),
    "pab_netvlad_forward_tc": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _L, _L, _P, _P]),
    "pab_afa_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "pab_gated_fc_workspace_bytes": (_SZ, [_I, _I, _I]),
    "pab_gated_fc_forward": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "pab_prepare_clouds": (_I, [_I, _I, _P, _I, C.POINTER(C.c_double), _I, _I, _P, _P, _P]),
    "pab_patch_triplets": (_I, [_I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, C.c_ulonglong, _I, _P, _P, _P, _P, _P]),
    "pab_afa_forward": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "pab_afa_tc_supported": (_I, [_I, _I, _I]),
    "pab_afa_tc_workspace_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "pab_afa_forward_tc": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "pab_tune_afa_tc": (None, [_I]),
    "pab_gated_fc_tc_supported": (_I, [_I, _I]),
    "pab_gated_fc_forward_tc": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
}


class PabError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PabError(
                f"{LIB_PATH} is missing: build it with `python -m patchaugnet_b200.build` "
                "(or __graft_entry__.build()); patchaugnet_b200 has no CPU or PyTorch fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def stream_ptr():
    """The caller's (PyTorch current) CUDA stream as a void*."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def check(rc, what):
    if rc == 0:
        return
    if rc == PAB_EINVAL:
        raise ValueError(f"{what}: argument outside the supported range (PAB_EINVAL)")
    raise PabError(f"{what}: CUDA error {rc}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PabError("patchaugnet_b200 ops run on CUDA tensors only (there is no CPU fallback)")
