"""`import pointops_cuda` shim: put <repo>/dropin on sys.path so the reference's libs/pointops/functions/pointops.py
(line 8) binds to the B200 kernels instead of the stock extension."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchaugnet_b200.pointops_cuda import *  # noqa: F401,F403,E402
