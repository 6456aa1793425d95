"""Fused eval-mode engine for PatchAugNet descriptor extraction (the north-star hot path).

Takes the parameters of a ``patchaugnet_b200.patch_aug_net.Network`` (same ``state_dict`` as the reference),
folds every eval-mode BatchNorm into its preceding weight once, and runs the whole forward
(``patch_aug_net.py:48-107`` in the reference) as ~25 launches of hand-written kernels through the C ABI:

    per SA level : FPS -> row gather (new_xyz) -> kNN(k) -> fused [group + centre-subtract + concat + SharedMLP + max_K]
    per FP level : fused [3-NN + inverse-distance weights] -> fused [interpolate + skip concat + SharedMLP]
    head         : NetVLAD level kernels (soft-assign + residual aggregation + intra-norm) -> AFA (attention, fc, BN, L2)

Activations are POINT-MAJOR (B, n, C) between kernels, so neighbour gathers read contiguous rows and NetVLAD consumes
the FP output without the reference's transpose+contiguous (loupe.py:192).  ``fp_features`` are returned as
(B, C, n, 1) *views* of those buffers — same shape and values as the reference, different strides.

Everything is launched on PyTorch's current stream, so the forward can be captured in a CUDA graph
(``capture_graph``) to remove per-launch host overhead.
"""
import ctypes as C

import torch

from . import _lib as L


def _fold_bn(bn):
    """eval-mode BatchNorm -> (scale, shift):  y = x*scale + shift"""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def _split_bf16(w):
    """w = hi + lo with hi = bf16(w), lo = bf16(w - hi): the two operand planes of the tensor-core path."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


class _Layers:
    """A folded SharedMLP as a ctypes array of pab_layer_t (keeps the device tensors alive).

    `extra_first` / `extra_last`: number of layer-0 input channels (the xyz part, <= 3) that sit before / after the
    block of channels eligible for the tensor-core path.  When every layer's remaining K and its N are multiples of
    64 the bf16 hi/lo weight planes are attached (pab_layer_t.w_hi / w_lo) and the C side picks the tcgen05 kernel.
    """

    def __init__(self, shared_mlp, device, extra_first=0, extra_last=0, precision="f32"):
        # precision "f32": bf16 hi/lo operand planes, three MMAs per product (the 1e-4 fp32 contract);
        #           "bf16": the rounded plane only (w_lo = NULL -> the kernels issue one MMA per product)
        self.precision = precision
        self.tensors = []
        blocks = list(shared_mlp.children())
        self.arr = (L.PabLayer * len(blocks))()
        self.spec = []
        folded = []
        for i, blk in enumerate(blocks):
            w = blk.conv.weight.detach().float().reshape(blk.conv.weight.shape[0], -1)   # (c_out, c_in)
            c_out, c_in = w.shape
            if hasattr(blk, "bn"):
                scale, shift = _fold_bn(blk.bn.bn)
            else:
                scale = torch.ones(c_out, device=w.device)
                shift = blk.conv.bias.detach().float() if blk.conv.bias is not None else torch.zeros(c_out, device=w.device)
            folded.append(((w * scale[:, None]).to(device), shift.to(device).contiguous(), c_in, c_out, hasattr(blk, "activation")))
        # tensor-core plan: "pre" = the first layer has a tiny input (<= 8 channels, SA1) and is evaluated by the kernel's
        # loader in fp32; every other layer needs K (zero-padded to a multiple of 64) and N in {32, 64k}
        pre = len(folded) >= 2 and folded[0][2] <= 8
        first_tc = 1 if pre else 0
        tc_ok = 1 <= len(folded) - first_tc <= 3
        for i, (wf, sh, c_in, c_out, act) in enumerate(folded):
            if i < first_tc:
                tc_ok = tc_ok and c_out % 16 == 0 and c_out <= 64
                continue
            kk = c_in - (extra_first + extra_last if i == 0 else 0)
            tc_ok = tc_ok and kk > 0 and (c_out == 32 or c_out % 64 == 0)
            if i == 0:
                tc_ok = tc_ok and kk % 64 == 0          # the gather loaders stage exactly K channels
        self.tensor_core = tc_ok
        for i, (wf, sh, c_in, c_out, act) in enumerate(folded):
            c_in_pad = (c_in + 3) // 4 * 4
            wt = torch.zeros(c_in_pad, c_out, dtype=torch.float32, device=device)
            wt[:c_in] = wf.t()
            self.tensors += [wt, sh]
            hi_p = lo_p = 0
            k0 = kpad = 0
            if tc_ok and i >= first_tc:
                k0 = extra_first if i == 0 else 0
                kk = c_in - (extra_first + extra_last if i == 0 else 0)
                kpad = (kk + 63) // 64 * 64
                n_extra = (extra_first + extra_last) if i == 0 else 0
                # layer 0 with extra (xyz) channels: one more 64-column chunk after the padded block holds their weights, in
                # channel order — the kernel feeds them to the tensor cores as an additional k-step (mlp_tc.cu)
                wk = torch.zeros(c_out, kpad + (64 if n_extra else 0), dtype=torch.float32, device=device)
                wk[:, :kk] = wf[:, k0:k0 + kk]
                if n_extra:
                    ex0 = 0 if extra_first else kk
                    wk[:, kpad:kpad + n_extra] = wf[:, ex0:ex0 + n_extra]
                hi, lo = _split_bf16(wk)
                self.tensors += [hi, lo]
                hi_p, lo_p = hi.data_ptr(), (lo.data_ptr() if precision == "f32" else 0)
            self.arr[i] = L.PabLayer(wt.data_ptr(), sh.data_ptr(), c_in, c_in_pad, c_out, 1 if act else 0, hi_p, lo_p, k0, kpad)
            self.spec.append((c_in, c_out))
        self.n = len(blocks)
        self.c_out = self.spec[-1][1]


class FusedPatchAugNet:
    """Eval-mode fused forward.  ``net`` is a patch_aug_net.Network on a CUDA device."""

    def __init__(self, net):
        self.net = net
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise L.PabError("FusedPatchAugNet needs the network on a CUDA device (there is no CPU fallback)")
        self._ws = {}
        self._graphs = {}
        self._events = None
        self._event_filter = None
        self._streams = None
        self.vlad_tensor_core = True
        self.fp_row_order = True        # FP modules walk their points in the Morton order of the level's spatial index
        self.dense_streams = 2          # dense kernels of consecutive batches alternate between two streams
        self.stream_priorities = (0, 0)
        self.stream_slots = 4           # forward_stream workspaces with the FPS stream: three stages in flight + one of slack, so the sampler
                                        # of batch i+3 does not wait for the dense kernels of batch i (41.2 k -> 41.8 k submaps/s)
        self.fps_stream = True          # forward_stream: the first level's FPS on a stream of its own (three-stage pipeline, three slots)
        self.stream_graphs = True       # forward_stream: geometry / dense launch sequences replayed as CUDA graphs (2 launches per batch)
        self._sgraphs = {}
        self.reserve_fps_sms = False    # forward_stream: persistent tensor-core kernels capped so that they leave the FPS CTAs' SMs alone
        self.stream_dynamic_tiles = True   # forward_stream: tensor-core CTAs draw their tiles from a counter (no cap needed: a CTA
                                           # whose SM is held by an FPS CTA starts late and takes fewer) — 40.6 k -> 42.3 k submaps/s
        self.tc_tune = 1                # pab_tune_tensor_core bits outside forward_stream
        self.fps_pack_from = 64         # forward_stream: launch sequences of at least this many clouds pack two clouds per FPS CTA
        self.fps_clouds_per_cta = 1     # forward_stream: clouds sharing one FPS CTA (2 = half the SMs held by the sampler)
        self.refold()

    # ---- weights -------------------------------------------------------------------------------------------------
    def refold(self):
        """(Re)build the folded inference weights from the module's current parameters."""
        net, dev = self.net, self.device
        bb = net.backbone
        self.sa = []
        for mod in bb.SA_modules:
            g = mod.groupers[0]
            self.sa.append(dict(npoint=mod.npoint, k=g.nsample, dilation=g.knn_dilation if g.radius is None else 1,
                                radius=g.radius, layers=_Layers(mod.mlps[0], dev, extra_first=3)))
        self.use_origin = bb.use_origin_pc_in_fp
        # FP_modules[0] concatenates the 3 raw xyz channels after the interpolated features (patch_aug_net.py:137, 359)
        self.fp = [_Layers(mod.mlp, dev, extra_last=3 if (i == 0 and self.use_origin) else 0)
                   for i, mod in enumerate(bb.FP_modules)]
        agg = net.aggregation
        self.vlad = []
        for v in agg.vlads:
            scale, shift = _fold_bn(v.bn1)
            wc = (v.cluster_weights.detach().float() * scale[None, :]).contiguous().to(dev)      # (C, K), bn1 folded
            w2 = v.cluster_weights2.detach().float()[0].contiguous().to(dev)                    # (C, K)
            K, Cf = v.cluster_size, v.feature_size
            Kp = (K + 15) // 16 * 16
            wct = torch.zeros(Kp, Cf, device=dev)
            wct[:K] = wc.t()
            hi, lo = _split_bf16(wct)                                                           # (Kp, C) K-major planes
            self.vlad.append(dict(K=K, C=Cf, wc=wc, shift=shift.contiguous().to(dev), w2=w2, wc_hi=hi, wc_lo=lo))
        self.c_out = agg.vlads[0].feature_size if agg.aggregation_type == 3 else agg.vlads[0].output_dim
        self._tail_shape = None
        self.sumK = sum(v["K"] for v in self.vlad)
        # aggregation_type 2 without context gating (the configured variant, patch_aug_net.yaml:9) has fused kernels; the other
        # variants (loupe.py:289-328) run the module's own tail on the fused NetVLAD outputs — a few tiny (B, 256, <=84) ops
        self.fused_tail = agg.aggregation_type == 2 and not agg.gating
        self._ws.clear()
        self._graphs.clear()
        self._sgraphs.clear()
        if not self.fused_tail:
            return
        afa = agg.afa
        self.w_att_t = afa.mlpa.mlps[0].weight.detach().float()[:, :, 0].t().contiguous().to(dev)   # (c_in, c_out)
        self.fc_wt = afa.fc.weight.detach().float().t().contiguous().to(dev)                        # (C*K, c_out)
        # tensor-core head (afa_tc.cu): both weights row-major over their inputs, split into bf16 hi/lo planes
        self.watt_planes = _split_bf16(afa.mlpa.mlps[0].weight.detach().float()[:, :, 0].contiguous().to(dev))   # (c_out', c_in)
        self.wfc_planes = _split_bf16(afa.fc.weight.detach().float().contiguous().to(dev))                       # (c_out, C*K)
        scale, shift = _fold_bn(afa.bn)
        self.fc_scale = scale.contiguous().to(dev)
        self.fc_shift = (afa.fc.bias.detach().float() * scale + shift).contiguous().to(dev)
        self.l2_norm = 1 if afa.l2_norm else 0
        self.c_out = afa.fc.weight.shape[0]

    # ---- workspace -----------------------------------------------------------------------------------------------
    def _workspace(self, B, N, slot=0):
        key = (B, N, slot)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        lib = L.lib()
        ws = dict(levels=[], fp=[])
        n = N
        for sa in self.sa:
            m = sa["npoint"]
            ws["levels"].append(dict(
                n=n, m=m,
                temp=torch.empty(B, n if n > 8192 else 1, **f32), cidx=torch.empty(B, m, **i32), new_xyz=torch.empty(B, m, 3, **f32),
                nbr=torch.empty(B, m, sa["k"], **i32), feat=torch.empty(B, m, sa["layers"].c_out, **f32),
                # Morton-sorted spatial index of the level's points (exact pruned kNN); small levels scan brute force
                index=(torch.empty(lib.pab_knn_index_bytes(B, n), dtype=torch.uint8, device=dev)
                       if 256 <= n <= 8192 and sa["k"] <= 64 else None)))
            n = m
        ns = [N] + [sa["npoint"] for sa in self.sa]              # points per level 0..3
        for li in range(len(self.fp)):                           # FP_modules[li] lifts level li+1 -> li
            ws["fp"].append(dict(idx=torch.empty(B, ns[li], 3, **i32), w=torch.empty(B, ns[li], 3, **f32),
                                 out=torch.empty(B, ns[li], self.fp[li].c_out, **f32)))
        ws["v"] = torch.empty(B, self.vlad[0]["C"], self.sumK, **f32)
        nbytes = max(lib.pab_netvlad_workspace_bytes(B, n_l, v["C"], v["K"])
                     for n_l, v in zip([ns[2], ns[1], ns[0]], self.vlad))
        if self.fused_tail:
            nbytes = max(nbytes, lib.pab_afa_workspace_bytes(B, self.vlad[0]["C"], self.sumK, self.c_out),
                         lib.pab_afa_tc_workspace_bytes(B, self.vlad[0]["C"], self.sumK, self.c_out))
        ws["scratch"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws["desc"] = torch.empty(B, self.c_out, **f32)
        self._ws[key] = ws
        return ws

    # ---- forward -------------------------------------------------------------------------------------------------
    def _runner(self):
        ev = self._events

        def run(stage, rc_fn):
            """launch one C-ABI call; with stage timing enabled, bracket it with CUDA events on the launch stream"""
            if ev is not None and (self._event_filter is None or stage in self._event_filter):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.check(rc_fn(), stage)
                e1.record()
                ev.setdefault(stage, []).append((e0, e1))
            else:
                L.check(rc_fn(), stage)
        return run

    def _launch_geo(self, xyz0, ws, first_knn_event=None, part=None):
        """Geometry of every level — FPS, centre gather, kNN, 3-NN weights.  Depends on xyz only, never on features,
        so it can run ahead of (and concurrently with) the dense part on another stream.
        part: None = everything; "fps0" = only the first level's sampler (one SM per cloud for 0.4 ms: forward_stream gives it
        a stream of its own); "rest" = everything after it."""
        lib, st, p, run = L.lib(), L.stream_ptr(), L.ptr, self._runner()
        B = xyz0.shape[0]
        xyz = xyz0
        for i, (sa, lv) in enumerate(zip(self.sa, ws["levels"])):
            n, m, k = lv["n"], lv["m"], sa["k"]
            if not (part == "rest" and i == 0):
                temp = None                      # register-resident FPS initialises its own 1e10 distances
                if n > 8192:                     # large-cloud fallback keeps them in global memory like the reference
                    temp = lv["temp"]
                    temp.fill_(1e10)
                run(f"fps{i}", lambda: lib.pab_furthestsampling(B, n, m, p(xyz), p(temp), p(lv["cidx"]), st))
            if part == "fps0":
                return
            run(f"gather{i}", lambda: lib.pab_gather_rows(B, n, m, 3, p(xyz), p(lv["cidx"]), p(lv["new_xyz"]), st))
            if lv["index"] is not None:
                run(f"index{i}", lambda: lib.pab_knn_build_index(B, n, p(xyz), p(lv["index"]), st))
            if sa["radius"] is not None:      # ball-query grouper (pointops.py:548-549); idx starts zeroed like BallQuery.forward
                lv["nbr"].zero_()
                run(f"knn{i}", lambda: lib.pab_ballquery(B, n, m, float(sa["radius"]), k, p(lv["new_xyz"]), p(xyz), p(lv["nbr"]), st))
            elif lv["index"] is not None:
                run(f"knn{i}", lambda: lib.pab_knnquery_indexed(B, n, m, k, p(lv["index"]), p(lv["new_xyz"]), p(lv["nbr"]), p(None), st))
            else:
                run(f"knn{i}", lambda: lib.pab_knnquery(B, n, m, k, p(xyz), p(lv["new_xyz"]), p(lv["nbr"]), p(None), st))
            if i == 0 and first_knn_event is not None:
                first_knn_event.record()      # the first SA module can start; deeper geometry overlaps with it
            xyz = lv["new_xyz"]
        xyzs = [xyz0] + [lv["new_xyz"] for lv in ws["levels"]]
        for li in range(len(self.fp) - 1, -1, -1):
            f = ws["fp"][li]
            unknown, known = xyzs[li], xyzs[li + 1]
            n, m = unknown.shape[1], known.shape[1]
            uidx = ws["levels"][li]["index"] if li < len(ws["levels"]) else None
            kidx = ws["levels"][li + 1]["index"] if li + 1 < len(ws["levels"]) else None
            if kidx is not None:              # spatial indices of both clouds were built for the kNN of their levels
                run(f"three_nn{li}", lambda: lib.pab_three_nn_weights_indexed(B, n, m, p(unknown), p(uidx), p(kidx),
                                                                              p(f["idx"]), p(f["w"]), st))
            else:
                run(f"three_nn{li}", lambda: lib.pab_three_nn_weights(B, n, m, p(unknown), p(known), p(f["idx"]), p(f["w"]), st))

    def _launch_dense(self, xyz0, ws, after_first_sa=None):
        """Feature path — fused SA modules, fused FP modules, NetVLAD levels, AFA head."""
        lib, st, p, run = L.lib(), L.stream_ptr(), L.ptr, self._runner()
        B = xyz0.shape[0]
        xyz, feat, c = xyz0, xyz0, 3
        for i, (sa, lv) in enumerate(zip(self.sa, ws["levels"])):
            n, m, k = lv["n"], lv["m"], sa["k"]
            run(f"sa{i}", lambda: lib.pab_sa_module_forward(B, n, m, k, k, c, p(xyz), p(feat), p(lv["cidx"]), p(lv["nbr"]),
                                                            sa["layers"].arr, sa["layers"].n, p(lv["feat"]), p(None), st))
            if i == 0 and after_first_sa is not None:
                after_first_sa()
            xyz, feat, c = lv["new_xyz"], lv["feat"], sa["layers"].c_out
        feats = [xyz0] + [lv["feat"] for lv in ws["levels"]]      # skip features per level (level 0 = raw xyz)
        ns = [xyz0.shape[1]] + [lv["m"] for lv in ws["levels"]]
        known_feat = feats[-1]
        for li in range(len(self.fp) - 1, -1, -1):                # FP_modules[-1] first (patch_aug_net.py:183-187)
            f = ws["fp"][li]
            n, m = ns[li], ns[li + 1]
            skip = feats[li]
            c_skip = skip.shape[2]
            if li == 0 and not self.use_origin:
                skip, c_skip = None, 0
            # rows in the Morton order of the level's spatial index (when it has one): neighbouring rows share their 3-NN rows
            order, stride = None, C.c_long(0)
            uindex = ws["levels"][li]["index"] if (self.fp_row_order and li < len(ws["levels"])) else None
            if uindex is not None:
                order = lib.pab_knn_index_order(n, p(uindex), C.byref(stride))
            run(f"fp{li}", lambda: lib.pab_fp_module_forward_ordered(B, n, m, known_feat.shape[2], c_skip, p(known_feat), p(skip),
                                                                     p(f["idx"]), p(f["w"]), C.c_void_p(order), stride.value,
                                                                     self.fp[li].arr, self.fp[li].n, p(f["out"]), st))
            known_feat = f["out"]
        # fp_features order of the reference: [l2 (128), l1 (1024), l0 (4096)] <-> vlads[0..2]
        fp_out = [ws["fp"][i]["out"] for i in range(len(self.fp) - 1, -1, -1)]
        v, koff = ws["v"], 0
        for i, (x, lvl) in enumerate(zip(fp_out, self.vlad)):
            dst = C.c_void_p(v.data_ptr() + 4 * koff)
            if self.vlad_tensor_core and lvl["C"] == 256:
                run(f"vlad{i}", lambda: lib.pab_netvlad_forward_tc(B, x.shape[1], lvl["C"], lvl["K"], p(x), p(lvl["wc_hi"]),
                                                                   p(lvl["wc_lo"]), p(lvl["shift"]), p(lvl["w2"]), dst, v.stride(0),
                                                                   v.stride(1), p(ws["scratch"]), st))
            else:
                run(f"vlad{i}", lambda: lib.pab_netvlad_forward(B, x.shape[1], lvl["C"], lvl["K"], p(x), p(lvl["wc"]), p(lvl["shift"]),
                                                                p(lvl["w2"]), dst, v.stride(0), v.stride(1), p(ws["scratch"]), st))
            koff += lvl["K"]
        if not self.fused_tail:
            per_level, koff = [], 0
            for lvl in self.vlad:
                per_level.append(v[:, :, koff:koff + lvl["K"]].contiguous())
                koff += lvl["K"]
            out = self.net.aggregation.aggregate(per_level)
            self._tail_shape = tuple(out.shape[1:])               # (c_out,) — or (c_out, 1): type 5 keeps AFA's last dim
            ws["desc"].copy_(out.reshape(B, -1))
            return fp_out
        if lib.pab_afa_tc_supported(self.vlad[0]["C"], self.sumK, self.c_out):
            run("afa", lambda: lib.pab_afa_forward_tc(B, self.vlad[0]["C"], self.sumK, self.c_out, p(v), p(self.watt_planes[0]),
                                                      p(self.watt_planes[1]), p(self.wfc_planes[0]), p(self.wfc_planes[1]),
                                                      p(self.fc_scale), p(self.fc_shift), self.l2_norm, p(ws["desc"]), p(ws["scratch"]), st))
        else:
            run("afa", lambda: lib.pab_afa_forward(B, self.vlad[0]["C"], self.sumK, self.c_out, p(v), p(self.w_att_t), p(self.fc_wt),
                                                   p(self.fc_scale), p(self.fc_shift), self.l2_norm, p(ws["desc"]), p(ws["scratch"]), st))
        return fp_out

    def _launch(self, xyz0, ws):
        """One forward on the current stream (also the body captured into a CUDA graph)."""
        self._launch_geo(xyz0, ws)
        return self._launch_dense(xyz0, ws)

    @torch.no_grad()
    def forward_stream(self, batches, out=None, ready_events=None, coalesce=0):
        """Throughput mode: descriptors of a sequence of equally shaped (B,N,3) / (B,1,N,3) CUDA batches.

        Submaps are independent, and the geometry of a batch (FPS is a serial m-step chain that occupies only B SMs)
        depends on nothing the dense path produces.  So batch i+1's geometry runs on a second stream while batch i's
        SharedMLP / NetVLAD kernels fill the rest of the machine; two workspaces ping-pong, events order the reuse.
        ``ready_events[i]`` (optional): a CUDA event batch i's geometry waits for — the host-to-device copy of that batch
        issued on a copy stream, so the upload of later batches overlaps the compute of earlier ones.
        ``coalesce`` (clouds, 0 = off): consecutive batches are concatenated into launch sequences of up to that many clouds.
        Submaps are independent and every kernel's arithmetic depends on the cloud only, so the descriptors are bit-identical; larger
        launches amortise per-launch set-up and the tails of the persistent kernels (41.5 k submaps/s with 32 clouds per
        sequence, 47.4 k with 64, 52.2 k with 128), at the price of latency and workspace (2.4 GB per slot at 128).
        Returns (len(batches)*B, c_out) descriptors on the device.
        """
        batches = list(batches)
        if not batches:
            return torch.empty(0, self.c_out, device=self.device)
        x0 = batches[0].squeeze(1) if batches[0].dim() == 4 else batches[0]
        B, N, _ = x0.shape
        if out is None:
            out = torch.empty(len(batches) * B, self.c_out, dtype=torch.float32, device=self.device)
        cur = torch.cuda.current_stream()
        g = int(coalesce) // B if coalesce else 0
        if g >= 2 and len(batches) >= g:
            # whole groups of g batches go through the pipeline as one batch each; a ragged remainder follows uncoalesced
            n_groups = len(batches) // g
            merged, events = [], []
            for gi in range(n_groups):
                grp = batches[gi * g:(gi + 1) * g]
                if ready_events is not None:
                    for ev in ready_events[gi * g:(gi + 1) * g]:
                        if ev is not None:
                            cur.wait_event(ev)
                merged.append(torch.cat([(x.squeeze(1) if x.dim() == 4 else x).float() for x in grp]))
                ev = torch.cuda.Event()
                ev.record(cur)
                events.append(ev)
            self.forward_stream(merged, out=out[:n_groups * g * B], ready_events=events)
            rest = batches[n_groups * g:]
            if rest:
                self.forward_stream(rest, out=out[n_groups * g * B:],
                                    ready_events=None if ready_events is None else ready_events[n_groups * g:])
            return out
        if self._streams is None:
            # stream_priorities: (geometry, dense) CUDA stream priorities, lower = scheduled first when SMs free up
            pg, pd = self.stream_priorities
            self._streams = (torch.cuda.Stream(device=self.device, priority=pg),) + tuple(
                torch.cuda.Stream(device=self.device, priority=pd) for _ in range(3))
        s_geo, dense_streams = self._streams[0], self._streams[1:1 + self.dense_streams]
        # Three-stage pipeline: the first level's FPS (a serial 1023-step chain on B SMs, 0.41 of the 0.76 ms geometry chain of a
        # batch) runs on its own stream, so FPS of batch i+2, the remaining geometry of batch i+1 and the dense kernels of batch i
        # overlap; with the whole geometry on one stream its chain was the period of the pipeline.
        s_fps = None
        if self.fps_stream:
            if getattr(self, "_fps_cuda_stream", None) is None:
                self._fps_cuda_stream = torch.cuda.Stream(device=self.device, priority=self.stream_priorities[0])
            s_fps = self._fps_cuda_stream
            s_fps.wait_stream(cur)
        n_slots = max(3, int(self.stream_slots)) if s_fps is not None else 2
        s_geo.wait_stream(cur)
        for sd in dense_streams:
            sd.wait_stream(cur)
        slots = [self._workspace(B, N, slot) for slot in range(n_slots)]
        # FPS holds one SM per cloud for a third of the step.  A persistent tensor-core kernel launched with one CTA per SM
        # would leave B of its CTAs waiting for those SMs and then run their static share of the tiles alone at the end
        # (measured: 28.3 k -> 31.2 k submaps/s at B = 32 with the cap; 120 instead of 116 CTAs is already slower than no cap).
        n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        if self.reserve_fps_sms:
            L.lib().pab_tune_fps_clouds_per_cta(self.fps_clouds_per_cta)
            cpc = max(1, L.lib().pab_fps_clouds_per_sm(N))          # what the sampler will really pack for this cloud size
            L.lib().pab_tune_tc_max_ctas(n_sm - (B + cpc - 1) // cpc if 0 < B <= n_sm // 2 else 0)
        else:
            # large launch sequences: two clouds per FPS CTA at 256 threads each (0.53 ms for the pair instead of 0.40 ms per
            # cloud: a third less SM-time for the sampler, its longer latency is hidden by the pipeline) — +3 % at 128 clouds
            # per sequence, -1 % at 32, hence the threshold
            pack = B >= self.fps_pack_from and self.fps_clouds_per_cta == 1
            L.lib().pab_tune_fps_threads(256 if pack else 0)
            L.lib().pab_tune_fps_clouds_per_cta(2 if pack else self.fps_clouds_per_cta)
            L.lib().pab_tune_tc_max_ctas(n_sm)                      # no cap; tells the small-CTA kernels that the SMs are shared
        if self.stream_dynamic_tiles:
            L.lib().pab_tune_tensor_core(self.tc_tune | 8)
        try:                                  # the tuning state above is process-global: restore it whatever happens
            fps_done = [None] * n_slots
            geo_done = [None] * n_slots
            dense_done = [None] * n_slots
            # The 30 launches of a batch are replayed as CUDA graphs (FPS, geometry, dense) per workspace slot: the host then issues
            # a copy + the graph launches + a few event operations per batch instead of ~30 ctypes calls, so the pipeline stays
            # GPU-bound when the host cores are contended (8 ranks on one box) — measured host time per batch 0.57 ms -> 0.1 ms.
            graphs = None
            if self.stream_graphs and self.fused_tail and self._events is None and len(batches) >= 4:
                graphs = self._capture_stream_graphs(B, N, slots, s_fps is not None)
            self.last_stream_used_graphs = graphs is not None
            for i, x in enumerate(batches):
                L.require_cuda(x)
                xyz0 = (x.squeeze(1) if x.dim() == 4 else x).contiguous().float()
                for sa in self.sa:                # keep the CPU RNG in lock-step with the reference (see forward)
                    if sa["dilation"] > 1:
                        torch.randperm(sa["k"])
                slot = i % n_slots
                ws = slots[slot]
                with torch.cuda.stream(s_fps if s_fps is not None else s_geo):
                    first = torch.cuda.current_stream()
                    if ready_events is not None and ready_events[i] is not None:
                        first.wait_event(ready_events[i])
                    if dense_done[slot] is not None:
                        first.wait_event(dense_done[slot])          # workspace (and the slot's static input) free again
                    if graphs is not None:
                        graphs[slot]["x"].copy_(xyz0, non_blocking=True)
                    if s_fps is not None:
                        if graphs is not None:
                            graphs[slot]["fps"].replay()
                        else:
                            self._launch_geo(xyz0, ws, part="fps0")
                        fps_done[slot] = torch.cuda.Event()
                        fps_done[slot].record()
                with torch.cuda.stream(s_geo):
                    if s_fps is not None:
                        s_geo.wait_event(fps_done[slot])
                    if graphs is not None:
                        graphs[slot]["geo"].replay()
                    else:
                        self._launch_geo(xyz0, ws, part="rest" if s_fps is not None else None)
                    geo_done[slot] = torch.cuda.Event()
                    geo_done[slot].record()
                s_dense = dense_streams[i % len(dense_streams)]   # alternate: the tail of one batch's kernels overlaps the next's
                with torch.cuda.stream(s_dense):
                    s_dense.wait_event(geo_done[slot])
                    if graphs is not None:
                        graphs[slot]["dense"].replay()
                    else:
                        self._launch_dense(xyz0, ws)
                    out[i * B:(i + 1) * B].copy_(ws["desc"], non_blocking=True)
                    dense_done[slot] = torch.cuda.Event()
                    dense_done[slot].record()
                xyz0.record_stream(s_geo)
                xyz0.record_stream(s_dense)
                if s_fps is not None:
                    xyz0.record_stream(s_fps)
            if s_fps is not None:
                cur.wait_stream(s_fps)
            for sd in dense_streams:
                cur.wait_stream(sd)
            cur.wait_stream(s_geo)
        finally:
            L.lib().pab_tune_tc_max_ctas(0)
            L.lib().pab_tune_fps_clouds_per_cta(1)
            if not self.reserve_fps_sms:
                L.lib().pab_tune_fps_threads(0)
            if self.stream_dynamic_tiles:
                L.lib().pab_tune_tensor_core(self.tc_tune)
        return out

    def _capture_stream_graphs(self, B, N, slots, split_fps=False):
        """FPS (optional) / geometry / dense launch sequences of every workspace slot as CUDA graphs over a static per-slot input
        buffer.  Captured with the current tuning state (CTA cap of the persistent kernels, FPS packing): the key holds it."""
        key = (B, N, L.lib().pab_fps_clouds_per_sm(N), self.reserve_fps_sms, self.fp_row_order, self.vlad_tensor_core,
               torch.cuda.get_device_properties(self.device).multi_processor_count, len(slots), split_fps,
               self.stream_dynamic_tiles, self.tc_tune)
        got = self._sgraphs.get(key)
        if got is not None:
            return got
        got = []
        cur = torch.cuda.current_stream()
        for ws in slots:
            x = torch.zeros(B, N, 3, dtype=torch.float32, device=self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                     # warm-up outside capture (function attributes, lazy state)
                self._launch_geo(x, ws)
                self._launch_dense(x, ws)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            g_fps, g_geo, g_dense = None, torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            if split_fps:
                g_fps = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_fps):
                    self._launch_geo(x, ws, part="fps0")
            with torch.cuda.graph(g_geo):
                self._launch_geo(x, ws, part="rest" if split_fps else None)
            with torch.cuda.graph(g_dense):
                self._launch_dense(x, ws)
            got.append(dict(x=x, fps=g_fps, geo=g_geo, dense=g_dense))
        self._sgraphs[key] = got
        return got

    # ---- per-stage timing and algorithmic work (bench.py roofline) ---------------------------------------------
    def enable_stage_timing(self, stages=None):
        """Record CUDA events around every stage (or only `stages`) of subsequent eager forwards."""
        self._events = {}
        self._event_filter = set(stages) if stages is not None else None

    def stage_times_ms(self):
        """{stage: [ms per recorded launch]} — call after torch.cuda.synchronize()."""
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in (self._events or {}).items()}

    def disable_stage_timing(self):
        self._events = None
        self._event_filter = None

    def stage_work(self, B, N):
        """Algorithmic work per launch of each stage: dense FLOPs (2*MAC) and compulsory HBM bytes."""
        ns = [N] + [sa["npoint"] for sa in self.sa]
        work = {}
        c = 3
        for i, sa in enumerate(self.sa):
            n, m, k = ns[i], ns[i + 1], sa["k"]
            macs = sum(ci * co for ci, co in sa["layers"].spec)
            c_out = sa["layers"].c_out
            work[f"fps{i}"] = dict(flops=0, bytes=B * (12 * n + 4 * m), units=B * n * (m - 1), unit="point-updates")
            work[f"gather{i}"] = dict(flops=0, bytes=B * (4 * m + 24 * m))
            work[f"index{i}"] = dict(flops=0, bytes=B * (12 * n + 16 * n + n // 2))       # read xyz, write sorted copy + perm + boxes
            work[f"knn{i}"] = dict(flops=0, bytes=B * (12 * n + 12 * m + 4 * m * k), units=B * n * m, unit="distance evals")
            work[f"sa{i}"] = dict(flops=2 * B * m * k * macs, bytes=B * (4 * c * n + 12 * n + 12 * m + 4 * m * k + 4 * c_out * m))
            c = c_out
        skip_c = [3] + [sa["layers"].c_out for sa in self.sa]
        known_c = skip_c[-1]
        for li in range(len(self.fp) - 1, -1, -1):
            n_u, n_k = ns[li], ns[li + 1]
            macs = sum(ci * co for ci, co in self.fp[li].spec)
            c_out = self.fp[li].c_out
            work[f"three_nn{li}"] = dict(flops=0, bytes=B * (12 * (n_u + n_k) + 24 * n_u), units=B * n_u * n_k, unit="distance evals")
            work[f"fp{li}"] = dict(flops=2 * B * n_u * macs,
                                   bytes=B * (4 * known_c * n_k + 4 * skip_c[li] * n_u + 24 * n_u + 4 * c_out * n_u))
            known_c = c_out
        for i, lvl in enumerate(self.vlad):
            n_l = ns[len(self.fp) - 1 - i]
            work[f"vlad{i}"] = dict(flops=2 * B * 2 * n_l * lvl["C"] * lvl["K"], bytes=B * 4 * (lvl["C"] * n_l + lvl["C"] * lvl["K"]))
        C_, K_ = self.vlad[0]["C"], self.sumK
        work["afa"] = dict(flops=2 * B * (C_ * C_ * K_ + C_ * K_ * self.c_out), bytes=4 * (C_ * K_ * self.c_out + C_ * C_ + B * C_ * K_ + B * self.c_out))
        return work

    @torch.no_grad()
    def forward(self, x, consume_rng=True, clone=True, return_feat=True):
        """x: (B,1,N,3) or (B,N,3) float32 CUDA -> (desc (B,256), fp_features [3 x (B,256,n,1)], center_idx_origin [3])."""
        L.require_cuda(x)
        xyz0 = x.squeeze(1) if x.dim() == 4 else x
        xyz0 = xyz0.contiguous().float()
        B, N, _ = xyz0.shape
        if consume_rng:
            # QueryAndGroup_Edge draws torch.randperm(nsample) on the CPU generator once per SA module when
            # knn_dilation > 1 (pointops.py:555).  The draw only permutes neighbour order (max-pool invariant);
            # it is repeated here so the global RNG stream stays in lock-step with the reference.
            for sa in self.sa:
                if sa["dilation"] > 1:
                    torch.randperm(sa["k"])
        ws = self._workspace(B, N)
        g = self._graphs.get((B, N))
        if g is not None:
            g["x"].copy_(xyz0)
            g["graph"].replay()
            fp_out = g["fp_out"]
        else:
            fp_out = self._launch(xyz0, ws)
        desc = ws["desc"]
        if self._tail_shape is not None:
            desc = desc.view(B, *self._tail_shape)
        if not return_feat:
            return desc.clone() if clone else desc
        cidx = [lv["cidx"] for lv in ws["levels"]]
        origin = [cidx[0]]
        for ci in cidx[1:]:
            origin.append(torch.gather(origin[-1], -1, ci.long()))       # patch_aug_net.py:169-177
        feats = [f.transpose(1, 2).unsqueeze(-1) for f in fp_out]
        if clone:   # detach results from the reusable workspace
            desc = desc.clone()
            feats = [f.clone(memory_format=torch.preserve_format) for f in feats]
            origin = [o.clone() for o in origin]
        return desc, feats, origin

    __call__ = forward

    def capture_graph(self, B, N):
        """Capture the (B,N) forward into a CUDA graph; later forwards of that shape replay it."""
        ws = self._workspace(B, N)
        x = torch.zeros(B, N, 3, dtype=torch.float32, device=self.device)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._launch(x, ws)          # warm-up (sets func attributes outside capture)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fp_out = self._launch(x, ws)
        self._graphs[(B, N)] = dict(graph=graph, x=x, fp_out=fp_out)
        return graph

    def launches_per_forward(self):
        """Kernels this library launches per forward (for bench.py's gpu_launches)."""
        n_index = max((sum(1 for lv in ws["levels"] if lv["index"] is not None) for ws in self._ws.values()), default=0)
        tail = 0
        if self.fused_tail:           # AFA head: attention logits, softmax, fc, finalize — the tensor-core fc kernel absorbs the softmax
            B = max((k[0] for k in self._ws), default=1)
            tc = L.lib().pab_afa_tc_supported(self.vlad[0]["C"], self.sumK, self.c_out)
            tail = 3 if (tc and min(B, 128) * self.sumK * 4 <= 32 * 1024) else 4
        return 4 * len(self.sa) + n_index + 2 * len(self.fp) + 2 * len(self.vlad) + tail
