"""CPU tier: the C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU),
the ctypes signature table covers the header, and the product has no route into the oracle."""
import ctypes
import os
import re

import pytest
import torch

import util
from patchaugnet_b200 import _lib as L
from patchaugnet_b200 import build as pab_build


def _declared_symbols():
    hdr = open(os.path.join(util.ROOT, "include", "patchaug_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pab_\w+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    pab_build.build()
    handle = ctypes.CDLL(L.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 36
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/patchaug_b200.h but not exported"
    assert handle.pab_version() == 1


def test_ctypes_table_covers_header():
    assert sorted(L.SIGNATURES) == _declared_symbols()


def test_library_targets_sm100a_only():
    out = os.popen(f"/usr/local/cuda/bin/cuobjdump -lelf {L.LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_ops_refuse_cpu_tensors_loudly():
    from patchaugnet_b200 import pointops
    xyz = torch.rand(1, 16, 3)
    with pytest.raises(L.PabError):
        pointops.furthestsampling(xyz, 4)
    with pytest.raises(L.PabError):
        pointops.knnquery(3, xyz, xyz)
    net = util.build_network()
    with pytest.raises(L.PabError):
        net.engine()                      # fused engine needs CUDA; no silent CPU path


def test_product_never_imports_the_oracle():
    pkg = os.path.join(util.ROOT, "patchaugnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle/" not in src, f


def test_dropin_shims_expose_reference_module_names():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "dropin"))
    try:
        m = importlib.import_module("pointops_cuda")
        # the 17 m.def names of libs/pointops/src/pointops_api.cpp:15-40
        for name in ["ballquery_cuda", "knnquery_cuda", "grouping_forward_cuda", "grouping_backward_cuda",
                     "grouping_int_forward_cuda", "gathering_forward_cuda", "gathering_backward_cuda", "furthestsampling_cuda",
                     "nearestneighbor_cuda", "interpolation_forward_cuda", "interpolation_backward_cuda", "labelstat_idx_cuda",
                     "labelstat_ballrange_cuda", "labelstat_and_ballquery_cuda", "featuredistribute_cuda",
                     "featuregather_forward_cuda", "featuregather_backward_cuda"]:
            assert callable(getattr(m, name))
        ch = importlib.import_module("chamfer")
        assert callable(ch.forward) and callable(ch.backward)
        em = importlib.import_module("emd")
        assert callable(em.forward) and callable(em.backward)
        kn = importlib.import_module("knn_cuda")
        assert kn.KNN(3, transpose_mode=True).k == 3
    finally:
        sys.path.remove(os.path.join(util.ROOT, "dropin"))
        for name in ("pointops_cuda", "chamfer", "emd", "knn_cuda"):
            sys.modules.pop(name, None)


def test_setup_py_installs_the_four_reference_import_names(tmp_path):
    """setup.py (the replacement of libs/pointops/setup.py:1-32 and of the three other extension installs): the built tree
    holds the package + the library and resolves `pointops_cuda`, `chamfer`, `emd`, `knn_cuda` without the repo on sys.path."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "lib")
    subprocess.check_call([sys.executable, "setup.py", "-q", "build", "--build-lib", out, "--build-temp", str(tmp_path / "tmp")],
                          cwd=root, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(out, "patchaugnet_b200", "libpatchaug_b200.so"))
    code = ("import pointops_cuda, chamfer, emd, knn_cuda, patchaugnet_b200, os;"
            "assert os.path.dirname(patchaugnet_b200.__file__).startswith(%r);"
            "names = [n for n in dir(pointops_cuda) if n.endswith('_cuda')]; assert len(names) == 17, names;"
            "assert callable(chamfer.forward) and callable(emd.backward) and knn_cuda.KNN(3).k == 3" % out)
    env = dict(os.environ, PYTHONPATH=out)
    subprocess.check_call([sys.executable, "-c", code], cwd=str(tmp_path), env=env)
    import shutil
    shutil.rmtree(os.path.join(root, "build"), ignore_errors=True)
    for d in (root, os.path.join(root, "dropin")):
        for e in os.listdir(d):
            if e.endswith(".egg-info"):
                shutil.rmtree(os.path.join(d, e), ignore_errors=True)
