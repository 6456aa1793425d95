"""Input pipeline in front of the descriptor path: .bin cloud -> offset -> normalise -> batch, with pinned
staging, asynchronous uploads double-buffered against the compute, and the arithmetic on the device.

Reference: ``SceneDataSet.make_descs`` / ``get_pcs`` / ``get_pc`` (``datasets/scene_dataset.py:494-523, 713-754``) load every
cloud with ``np.fromfile``, subtract ``global_offset``, call ``normalize_point_cloud``
(``utils/loading_pointclouds.py:51-63``) on the host, stack the batch and upload it synchronously before each forward.
Here the file bytes go straight into a pinned staging buffer, one upload per batch runs on a copy stream, and
``pab_prepare_clouds`` (csrc/prepare.cu) does offset / centre / scale / cast for the whole batch in one launch.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def load_pc_file(path, dtype=np.float64):
    """``utils/loading_pointclouds.py:14-24`` (3-D branch): the raw (N, 3) array of a ``.bin`` cloud."""
    return np.fromfile(path, dtype=dtype).reshape([-1, 3])


def prepare_batch(raw, global_offset=None, normalize=False, zoom=True, return_norm_meta=False):
    """raw (B, N, 3) float64 / float32 CUDA tensor -> the (B, 1, N, 3) float32 batch ``Network.forward`` takes
    (``get_pc`` + ``normalize_point_cloud`` of every cloud, one launch).  With ``return_norm_meta`` also the list of
    ``{'scale', 'trans'}`` dicts the reference keeps per cloud."""
    L.require_cuda(raw)
    if raw.dtype not in (torch.float64, torch.float32) or raw.dim() != 3 or raw.shape[2] != 3:
        raise ValueError("raw clouds must be (B, N, 3) float64 or float32")
    raw = raw.contiguous()
    B, N, _ = raw.shape
    out = torch.empty(B, 1, N, 3, dtype=torch.float32, device=raw.device)
    meta = torch.empty(B, 4, dtype=torch.float64, device=raw.device) if return_norm_meta else None
    off = (C.c_double * 3)(*(np.asarray(global_offset, dtype=np.float64).reshape(-1)[:3] if global_offset is not None else (0.0, 0.0, 0.0)))
    L.check(L.lib().pab_prepare_clouds(B, N, L.ptr(raw), 1 if raw.dtype == torch.float64 else 0, off, 1 if normalize else 0,
                                       1 if zoom else 0, L.ptr(out), L.ptr(meta), L.stream_ptr()), "pab_prepare_clouds")
    if not return_norm_meta:
        return out
    m = meta.cpu().numpy()
    return out, [{"scale": float(r[0]), "trans": r[1:4].copy()} for r in m]


class CloudFeeder:
    """Prepared device batches from ``.bin`` files (or in-memory raw arrays), uploads overlapped with the consumer's compute.

    Iterating yields ``(batch (b,1,N,3) float32 CUDA, ready_event)``: the batch is produced on the feeder's copy stream
    (upload + ``pab_prepare_clouds``); a consumer stream waits for ``ready_event`` before it touches the batch
    (``FusedPatchAugNet.forward_stream(..., ready_events=...)``).  Pinned staging buffers rotate (``depth``), a buffer is
    refilled only after its upload has completed.
    """

    def __init__(self, sources, batch_size, num_points=4096, dtype=np.float64, global_offset=None, normalize=False, zoom=True,
                 device="cuda", depth=3):
        self.sources, self.batch_size, self.n = list(sources), batch_size, num_points
        self.dtype, self.offset, self.normalize, self.zoom = np.dtype(dtype), global_offset, normalize, zoom
        self.device = torch.device(device)
        tdtype = torch.float64 if self.dtype == np.float64 else torch.float32
        self.staging = [torch.empty(batch_size, num_points, 3, dtype=tdtype).pin_memory() for _ in range(depth)]
        self.uploaded = [None] * depth
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return (len(self.sources) + self.batch_size - 1) // self.batch_size

    def _read(self, src, dst):
        pc = load_pc_file(src, self.dtype) if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__") else np.asarray(src)
        if pc.shape != (self.n, 3):
            raise ValueError(f"cloud has shape {pc.shape}, expected ({self.n}, 3)")
        dst.copy_(torch.from_numpy(np.ascontiguousarray(pc, dtype=self.dtype)))

    def __iter__(self):
        for bi in range(len(self)):
            slot = bi % len(self.staging)
            if self.uploaded[slot] is not None:
                self.uploaded[slot].synchronize()               # the staging buffer's previous upload has left the host
            srcs = self.sources[bi * self.batch_size:(bi + 1) * self.batch_size]
            buf = self.staging[slot][:len(srcs)]
            for j, src in enumerate(srcs):
                self._read(src, buf[j])
            with torch.cuda.stream(self.stream):
                raw = buf.to(self.device, non_blocking=True)
                self.uploaded[slot] = torch.cuda.Event()
                self.uploaded[slot].record(self.stream)
                batch = prepare_batch(raw, self.offset, self.normalize, self.zoom)
                ready = torch.cuda.Event()
                ready.record(self.stream)
            batch.record_stream(torch.cuda.current_stream())
            yield batch, ready


def make_descs(net, sources, batch_size=32, num_points=4096, dtype=np.float64, global_offset=None, normalize=False, zoom=True,
               device="cuda", super_chunk=16):
    """Global descriptors of a list of ``.bin`` clouds — the generic branch of ``SceneDataSet.make_descs``
    (``scene_dataset.py:666-708``) — through the feeder and the fused engine's throughput mode.  Returns (M, 256) CUDA."""
    feeder = CloudFeeder(sources, batch_size, num_points, dtype, global_offset, normalize, zoom, device)
    engine = net.engine() if (hasattr(net, "fusable") and not net.training and net.fusable()) else None
    outs, pend_b, pend_e = [], [], []

    def flush():
        if not pend_b:
            return
        full = [b for b in pend_b if b.shape[0] == batch_size]
        if engine is not None and full:
            outs.append(engine.forward_stream(full, ready_events=pend_e[:len(full)]))
        for b, e in list(zip(pend_b, pend_e))[len(full) if engine is not None else 0:]:
            torch.cuda.current_stream().wait_event(e)
            d = net(b, return_feat=False) if engine is not None else net(b)
            outs.append((d[0] if isinstance(d, (tuple, list)) else d).reshape(b.shape[0], -1).clone())
        pend_b.clear(); pend_e.clear()

    with torch.no_grad():
        for batch, ready in feeder:
            pend_b.append(batch); pend_e.append(ready)
            if len(pend_b) == super_chunk:
                flush()
        flush()
    return torch.cat(outs) if outs else torch.empty(0, 256, device=device)
