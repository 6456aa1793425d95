"""`import emd` shim (libs/emd_module/emd_module.py:26): forward / backward over libpatchaug_b200.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchaugnet_b200.emd_module import forward, backward  # noqa: F401,E402
