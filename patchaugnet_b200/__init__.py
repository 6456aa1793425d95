"""patchaugnet_b200 — B200-native (sm_100a) descriptor-extraction-and-retrieval hot path of WHU-USI3DV/PatchAugNet.

Host-side mirrors of the reference's operator / nn.Module API over hand-written CUDA kernels reached through the
C ABI in ``include/patchaug_b200.h`` (``libpatchaug_b200.so``, built by ``patchaugnet_b200.build``).
There is no CPU, PyTorch-eager or Triton fallback: ops raise if the library is missing or a tensor is not on CUDA.
"""
__version__ = "0.1.0"
