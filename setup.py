"""Packaging of the B200 path behind the reference's four extension names.

    python setup.py install          # or: pip install --no-build-isolation .

The reference installs its native code with `python setup.py install` in libs/pointops, libs/chamfer_dist and libs/emd_module
(README.md:37-43; libs/pointops/setup.py:1-32 builds the pybind module `pointops_cuda`) and JIT-builds `knn_cuda` at import
(libs/KNN_CUDA/knn_cuda/__init__.py:10-38).  This setup.py replaces all four: it compiles libpatchaug_b200.so for sm_100a
(patchaugnet_b200/build.py: nvcc -gencode arch=compute_100a,code=sm_100a) and installs

    patchaugnet_b200/     the package (C-ABI library, ctypes bindings, module mirrors, fused engines)
    pointops_cuda.py      -> libs/pointops/functions/pointops.py:8   `import pointops_cuda`
    chamfer.py            -> libs/chamfer_dist/__init__.py:10         `import chamfer`
    emd.py                -> libs/emd_module/emd_module.py:26         `import emd`
    knn_cuda/             -> utils/train_util.py:14                   `from knn_cuda import KNN`

so the reference's Python runs unchanged on the new kernels (tests/test_reference_python_gpu.py executes exactly that).
"""
import os
import sys

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class BuildWithCuda(build_py):
    def run(self):
        sys.path.insert(0, HERE)
        from patchaugnet_b200 import build as pab_build
        pab_build.build()                     # in-tree libpatchaug_b200.so (no-op when up to date)
        super().run()


setup(
    name="patchaugnet_b200",
    version="0.2.0",
    description="B200-native descriptor-extraction-and-retrieval path of PatchAugNet behind the reference's operator API",
    packages=["patchaugnet_b200", "knn_cuda"],
    package_dir={"": "dropin", "patchaugnet_b200": "patchaugnet_b200"},
    py_modules=["pointops_cuda", "chamfer", "emd"],
    package_data={"patchaugnet_b200": ["libpatchaug_b200.so"]},
    cmdclass={"build_py": BuildWithCuda},
    python_requires=">=3.9",
)
