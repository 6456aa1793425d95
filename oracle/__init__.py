"""CPU oracle for the PatchAugNet descriptor-extraction hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``patchaugnet_b200``) never does, and fails loudly when its
CUDA library is missing instead of falling back to anything in here.

* ``oracle.ops``   — ctypes bindings of ``liboracle.so`` (``pointops_oracle.c`` / ``emd_oracle.c``): the C
  restatement of the reference's CUDA kernels, citing reference file:line per function.
* ``oracle.model`` — torch-CPU restatement of the reference's nn.Module forwards, built on ``oracle.ops``.
* ``oracle.refgpu`` — ctypes bindings of ``oracle/_ref/*.so``: the reference's OWN kernels compiled from
  /root/reference for sm_100 (GPU box only); used to pin the restatement.
"""
