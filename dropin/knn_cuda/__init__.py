"""`from knn_cuda import KNN` shim (utils/train_util.py:14): no import-time JIT build, no CUDA assert at import."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from patchaugnet_b200.knn_cuda import KNN, knn, _knn, _T, __version__  # noqa: F401,E402
