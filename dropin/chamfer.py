"""`import chamfer` shim (libs/chamfer_dist/__init__.py:10): forward / backward over libpatchaug_b200.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchaugnet_b200.chamfer_dist import forward, backward  # noqa: F401,E402
