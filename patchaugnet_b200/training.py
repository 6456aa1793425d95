"""Training step of PatchAugNet (BASELINE.json configs[4]) — the reference's ``run_model`` + loss assembly + optimiser step
(``place_recognition/train_place_recognition.py:142-169, 274-306, 386-392``) restated on this repo's modules, data-parallel
over anchors with ``torch.distributed`` (one process per GPU, NCCL).

The reference trains with single-process ``nn.DataParallel`` (``train_place_recognition.py:546-548``): the batch of tuples
is scattered over the GPUs, BatchNorm statistics stay per replica, gradients are summed on GPU 0.  The equivalent here is
``DistributedDataParallel`` over the anchors: every rank gets ``anchors / world`` tuples of 18 clouds (1 query, 2 positives,
14 negatives, 1 other negative; ``patch_aug_net.yaml:61-62``), BatchNorm stays per rank, and the one collective is DDP's
bucketed all-reduce of the 13.47 M fp32 gradients (53.9 MB) over NCCL / NVLink, overlapped with the backward pass.  Every
loss term is a mean over anchors / patches, so the average of the per-rank gradients is the gradient of the global batch.

``run_model`` as shipped cannot run (``Network.forward`` defaults to ``return_feat=True`` and returns
a 3-tuple that ``run_model`` unpacks into two names); the step below calls ``model(feed, nn_dict, return_feat=False)``, which
is the call the script intends.

Backward kernels: gathering / grouping / interpolation scatter their gradients deterministically (``csrc/scatter.cu``: an
inverted index per cloud, ordered sums, no atomics), the patch chamfer backward is a per-patch gather (``csrc/chamfer.cu``);
the dense layers run on PyTorch / cuDNN in train mode (batch statistics) — the fused tcgen05 kernels are eval-only.
"""
import torch
import torch.distributed as dist

from . import losses

CLOUDS_PER_ANCHOR = 18          # 1 query + TRAIN_POSITIVES_PER_QUERY (2) + TRAIN_NEGATIVES_PER_QUERY (14) + 1 other negative


def make_nn_dict(n_anchors, positives_per_query=2, clouds_per_anchor=CLOUDS_PER_ANCHOR, pairs=None):
    """Keys of the reference's ``nn_dict``: (anchor cloud, positive cloud) index pairs inside the flattened feed tensor,
    ``(j*18 + 0, j*18 + p)`` for p in 1..positives (datasets/scene_dataset.py:293-296, train_place_recognition.py:260-265).
    Values are the overlap index pairs of the two clouds (only the a2b term reads them; ``pairs`` or an empty list)."""
    out = {}
    for j in range(n_anchors):
        for p in range(1, positives_per_query + 1):
            out[(j * clouds_per_anchor, j * clouds_per_anchor + p)] = pairs if pairs is not None else []
    return out


def split_descriptors(desc, n_anchors, positives=2, negatives=14):
    """``run_model``'s split (train_place_recognition.py:165-169): (A*18, D) -> q (A,1,D), pos (A,2,D), neg (A,14,D), other (A,1,D)."""
    d = desc.view(n_anchors, -1, desc.shape[-1])
    return torch.split(d, [1, positives, negatives, 1], dim=1)


def assemble_loss(desc, patch_recon, n_anchors, margin_1=0.5, margin_2=0.2, lazy=True, use_min=False, ignore_zero=False,
                  weight_place=1.0, weight_patch_recon=0.25):
    """Place-recognition quadruplet loss + patch-reconstruction chamfer loss with the shipped weights
    (patch_aug_net.yaml:9-12, 76-83; train_place_recognition.py:279-306, 386-390)."""
    q, pos, neg, other = split_descriptors(desc, n_anchors)
    terms = {"place_recognition": losses.quadruplet_loss(q, pos, neg, other, margin_1, margin_2, use_min=use_min, lazy=lazy,
                                                         ignore_zero_loss=ignore_zero)}
    total = weight_place * terms["place_recognition"]
    if patch_recon is not None and len(patch_recon["reconstructed_patches"]):
        terms["patch_recon_a2a"] = losses.patch_chamfer_loss(patch_recon["origin_patches"], patch_recon["reconstructed_patches"])
        total = total + weight_patch_recon * terms["patch_recon_a2a"]
    return total, terms


class TrainStep:
    """One optimisation step over this rank's tuples.  ``model``: a ``patch_aug_net.Network`` (or its DDP wrapper) in train mode."""

    def __init__(self, model, optimizer, n_anchors, use_patch_recon=True):
        self.model, self.optimizer, self.n_anchors, self.use_patch_recon = model, optimizer, n_anchors, use_patch_recon
        self.nn_dict = make_nn_dict(n_anchors) if use_patch_recon else None

    def __call__(self, feed):
        """feed: (n_anchors*18, 1, N, 3) float32 on the model's device.  Returns (loss, dict of detached loss terms)."""
        feed = feed.detach().requires_grad_(True)                 # train_place_recognition.py:150
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(feed, self.nn_dict, return_feat=False)
        desc, recon = out if isinstance(out, tuple) else (out, None)
        loss, terms = assemble_loss(desc, recon, self.n_anchors)
        loss.backward()
        self.optimizer.step()
        return loss.detach(), {k: v.detach() for k, v in terms.items()}


class GraphedTrainStep(TrainStep):
    """``TrainStep`` captured ONCE into a CUDA graph (forward, losses, backward, optimiser step) and replayed: the ~1500 kernel
    launches and the Python of a step cost the host ~55 ms on an idle box and several times that on a contended one, a graph
    launch costs microseconds, so the step becomes GPU-bound whatever the host does (measured 86.5 ms vs 101-115 ms eager for
    288 clouds).  Also works on a DistributedDataParallel model (NCCL all-reduce captured with the backward: 86.8 ms per step on
    two GPUs, weights in sync) when the wrapper is built by ``build_ddp(..., for_graph=True)`` and ``warmup >= 11``; release
    the step (``step.release()``) BEFORE ``destroy_process_group`` — tearing NCCL down under a live captured graph hangs.  The
    optimiser must be created with ``capturable=True``.  Frozen in the graph: the tuple layout (``nn_dict``), the neighbour ORDER drawn by
    the groupers' ``torch.randperm`` at the last eager step (order-invariant downstream), the learning rate."""

    def __init__(self, model, optimizer, n_anchors, use_patch_recon=True, warmup=3):
        super().__init__(model, optimizer, n_anchors, use_patch_recon)
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise ValueError("GraphedTrainStep needs an optimizer created with capturable=True")
        self._graph, self._feed, self._out, self._warmup = None, None, None, warmup

    def _body(self):
        x = self._feed.detach().requires_grad_(True)
        out = self.model(x, self.nn_dict, return_feat=False)
        desc, recon = out if isinstance(out, tuple) else (out, None)
        loss, terms = assemble_loss(desc, recon, self.n_anchors)
        loss.backward()
        self.optimizer.step()
        return loss.detach(), {k: v.detach() for k, v in terms.items()}

    def __call__(self, feed):
        if self._graph is None:
            self._feed = feed.detach().clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=feed.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                       # eager warm-up on a side stream (allocator, cuDNN plans, Adam state)
                for _ in range(self._warmup):
                    self.optimizer.zero_grad(set_to_none=True)
                    self._body()
            cur.wait_stream(side)
            torch.cuda.synchronize(feed.device)
            self._graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(self._graph):
                self._out = self._body()
            self._graph.replay()                                # capture records, it does not execute: this call's own step
        else:
            self._feed.copy_(feed)
            self._graph.replay()
        return self._out

    def release(self):
        """Drop the captured graph and its memory pool (call before destroying the process group of a DDP model)."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph.reset()
        self._graph, self._out = None, None


def build_ddp(model, device, for_graph=False):
    """Wrap ``model`` for data-parallel training when a process group is initialised (one process per GPU).
    ``for_graph``: the wrapper CUDA-graph capture needs (built on a side stream, static graph, every parameter used)."""
    if for_graph and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], output_device=device.index,
                                                            gradient_as_bucket_view=True, static_graph=True)
        torch.cuda.current_stream().wait_stream(side)
        return ddp
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        ddp_kwargs = dict(device_ids=[device.index], output_device=device.index) if device.type == "cuda" else {}
        # the decoder / unused NetVLAD parameters (hidden1_weights, bn2, mlpa.trans_conv ...) receive no gradient in a step
        return torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True, gradient_as_bucket_view=True, **ddp_kwargs)
    return model


def grad_bytes(model):
    return sum(p.numel() * p.element_size() for p in model.parameters() if p.requires_grad)
