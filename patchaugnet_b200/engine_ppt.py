"""Fused eval-mode forward of PPT-Net (``patchaugnet_b200.pptnet.Network``) on the hand-written kernels.

Same building blocks as ``engine.FusedPatchAugNet`` — point-major features, BatchNorm folded into the weights, one C-ABI
call per module — arranged as the reference's ``pptnet.py:65-134`` backbone:

    level i (4096 -> 1024 -> 256 -> 64 -> 16 points):
        FPS -> centre gather -> [Morton-chunk index] -> kNN -> fused SA module (gather, edge features, SharedMLP, max over K,
        ``mlp_tc_kernel`` / ``mlp_kernel``) -> fused ``SA_Layer`` self-attention (``attention.cu``)
    4 feature-propagation modules (3-NN weights + fused FP module), deepest first
    NetVLAD on the four pyramid levels (64 / 256 / 1024 / 4096 points, 1 / 4 / 16 / 64 clusters) written straight into the
    reference's flattened (B, C*K) layout, then ``hidden_weights`` -> bn2 -> context gating -> L2 as one split-K fc kernel +
    one finalize kernel (``pab_gated_fc_forward``, row a14)

Numerics follow the module mirror (``pptnet.py``) to fp32 round-off; ``tests/test_pptnet_gpu.py`` checks both against the
reference golden vectors and against each other.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib as L
from . import attention
from .engine import _Layers, _fold_bn, _split_bf16


class FusedPPTNet:
    def __init__(self, net, precision="f32"):
        """precision "f32": the reference's fp32 contract (bf16 hi/lo tensor-core operands, descriptors within 1e-4);
        "bf16": BASELINE.json configs[2] — every tensor-core operand rounded to bf16 once, fp32 accumulation, geometry
        (FPS / kNN / 3-NN indices and weights) unchanged and still bit-exact."""
        self.net = net
        self.precision = precision
        # per-part arithmetic.  "bf16" mode: FP modules, NetVLAD and the attention passes use plain bf16 operands (one MMA per
        # product); the SA modules keep the hi/lo planes — their inputs are coordinate / feature DIFFERENCES of neighbouring
        # points, and rounding those to 8 mantissa bits alone costs the descriptor 4e-3 of cosine (measured, profiles/r02_results.md),
        # which would break the configuration's parity definition (cosine >= 0.999 against the fp32 forward)
        self.fp_precision = self.vlad_precision = precision
        self.sa_precision = "f32"
        self.attention_precision = 2 if precision == "f32" else 1     # tensor-core attention: hi/lo planes or plain bf16 (0 = SIMT)
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise L.PabError("FusedPPTNet needs the network on a CUDA device (there is no CPU fallback)")
        self._ws = {}
        self.refold()

    # ---- weights -------------------------------------------------------------------------------------------------
    def refold(self):
        net, dev = self.net, self.device
        bb = net.backbone
        self.sa = []
        for mod in bb.SA_modules:
            g = mod.groupers[0]
            self.sa.append(dict(npoint=mod.npoint, k=g.nsample, layers=_Layers(mod.mlps[0], dev, extra_first=3, precision=self.sa_precision),
                                att=attention._fold(mod.sas[0], dev)))
        # FP_modules[0] takes the raw xyz (3 channels) as its skip input (pptnet.py:83-90: l_features[0] = xyz^T)
        self.fp = [_Layers(mod.mlp, dev, extra_last=3 if i == 0 else 0, precision=self.fp_precision) for i, mod in enumerate(bb.FP_modules)]
        agg = net.aggregation
        self.vlad = []
        for i in range(4):
            v = getattr(agg, f"vlad{i}")
            scale, shift = _fold_bn(v.bn1)
            wc = (v.cluster_weights.detach().float() * scale[None, :]).contiguous().to(dev)      # (C, K), bn1 folded
            w2 = v.cluster_weights2.detach().float()[0].contiguous().to(dev)                    # (C, K)
            K, Cf = v.cluster_size, v.feature_size
            Kp = (K + 15) // 16 * 16
            wct = torch.zeros(Kp, Cf, device=dev)
            wct[:K] = wc.t()
            hi, lo = _split_bf16(wct)
            lvl = dict(K=K, C=Cf, n=v.max_samples, wc=wc, shift=shift.contiguous().to(dev), w2=w2, wc_hi=hi, wc_lo=lo)
            if K % 4:
                # the NetVLAD kernels take cluster counts in multiples of four: pad with clusters that can never be assigned
                # (logit shift -1e30 -> softmax weight exactly 0), run on the padded problem and keep the real columns
                K4 = (K + 3) // 4 * 4
                wc4 = torch.zeros(Cf, K4, device=dev); wc4[:, :K] = wc
                sh4 = torch.full((K4,), -1e30, device=dev); sh4[:K] = lvl["shift"]
                w24 = torch.zeros(Cf, K4, device=dev); w24[:, :K] = w2
                lvl.update(K4=K4, wc4=wc4.contiguous(), shift4=sh4, w24=w24.contiguous())
            self.vlad.append(lvl)
        self.flat = sum(v["K"] * v["C"] for v in self.vlad)
        # head: hidden_weights -> bn2 -> context gating (bn1 or biases folded) -> L2, one split-K fc + one finalize kernel
        self.fc_wt = agg.hidden_weights.detach().float().contiguous().to(dev)                  # (flat, c_out)
        self.c_out = self.fc_wt.shape[1]
        self.fc_planes = _split_bf16(self.fc_wt.t().contiguous())                              # (c_out, flat) hi / lo for the tensor-core head
        sc, sh = _fold_bn(agg.bn2)
        self.fc_scale, self.fc_shift = sc.contiguous().to(dev), sh.contiguous().to(dev)
        self.gate = None
        if agg.gating:
            cg = agg.context_gating
            gw = cg.gating_weights.detach().float().contiguous().to(dev)                       # (c_out, c_out): gates = x @ G
            if cg.add_batch_norm:
                gs, gb = _fold_bn(cg.bn1)
            else:
                gs, gb = torch.ones(self.c_out, device=dev), cg.gating_biases.detach().float()
            self.gate = (gw, gs.contiguous().to(dev), gb.contiguous().to(dev))
        self._ws.clear()

    # ---- workspace -----------------------------------------------------------------------------------------------
    def _workspace(self, B, N, slot=0):
        key = (B, N, slot)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev, lib = self.device, L.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        ws = dict(levels=[], fp=[])
        n = N
        att_bytes = 0
        for sa in self.sa:
            m, c_out = sa["npoint"], sa["layers"].c_out
            ws["levels"].append(dict(
                n=n, m=m, cidx=torch.empty(B, m, **i32), new_xyz=torch.empty(B, m, 3, **f32), nbr=torch.empty(B, m, sa["k"], **i32),
                pooled=torch.empty(B, m, c_out, **f32), feat=torch.empty(B, m, c_out, **f32),
                temp=torch.empty(B, n if n > 8192 else 1, **f32),
                index=(torch.empty(lib.pab_knn_index_bytes(B, n), dtype=torch.uint8, device=dev)
                       if 256 <= n <= 8192 and sa["k"] <= 64 else None)))
            att_bytes = max(att_bytes, lib.pab_sa_layer_workspace_bytes(B, m, c_out))
            n = m
        ns = [N] + [sa["npoint"] for sa in self.sa]
        for li in range(len(self.fp)):
            ws["fp"].append(dict(idx=torch.empty(B, ns[li], 3, **i32), w=torch.empty(B, ns[li], 3, **f32),
                                 out=torch.empty(B, ns[li], self.fp[li].c_out, **f32)))
        ws["vlad"] = torch.empty(B, self.flat, **f32)
        ws["vlad_pad"] = torch.empty(B, self.vlad[0]["C"], 4, **f32)
        ws["desc"] = torch.empty(B, self.c_out, **f32)
        nbytes = max(lib.pab_netvlad_workspace_bytes(B, v["n"], v["C"], max(v["K"], 4)) for v in self.vlad)
        nbytes = max(nbytes, lib.pab_gated_fc_workspace_bytes(B, self.flat, self.c_out))
        ws["scratch"] = torch.empty(max(nbytes, att_bytes), dtype=torch.uint8, device=dev)
        self._ws[key] = ws
        return ws

    # ---- forward -------------------------------------------------------------------------------------------------
    def _launch_geo(self, xyz0, ws):
        """Geometry of every level — FPS, centre gather, spatial index, kNN, 3-NN weights.  Depends on xyz only."""
        lib, st, p, chk = L.lib(), L.stream_ptr(), L.ptr, L.check
        B = xyz0.shape[0]
        xyz = xyz0
        for sa, lv in zip(self.sa, ws["levels"]):
            n, m, k = lv["n"], lv["m"], sa["k"]
            temp = None
            if n > 8192:
                temp = lv["temp"]
                temp.fill_(1e10)
            chk(lib.pab_furthestsampling(B, n, m, p(xyz), p(temp), p(lv["cidx"]), st), "fps")
            chk(lib.pab_gather_rows(B, n, m, 3, p(xyz), p(lv["cidx"]), p(lv["new_xyz"]), st), "gather")
            if lv["index"] is not None:
                chk(lib.pab_knn_build_index(B, n, p(xyz), p(lv["index"]), st), "index")
                chk(lib.pab_knnquery_indexed(B, n, m, k, p(lv["index"]), p(lv["new_xyz"]), p(lv["nbr"]), p(None), st), "knn")
            else:
                chk(lib.pab_knnquery(B, n, m, k, p(xyz), p(lv["new_xyz"]), p(lv["nbr"]), p(None), st), "knn")
            xyz = lv["new_xyz"]
        xyzs = [xyz0] + [lv["new_xyz"] for lv in ws["levels"]]
        ns = [xyz0.shape[1]] + [lv["m"] for lv in ws["levels"]]
        for li in range(len(self.fp) - 1, -1, -1):
            f = ws["fp"][li]
            n, m = ns[li], ns[li + 1]
            uidx = ws["levels"][li]["index"] if li < len(ws["levels"]) else None
            kidx = ws["levels"][li + 1]["index"] if li + 1 < len(ws["levels"]) else None
            if kidx is not None:
                chk(lib.pab_three_nn_weights_indexed(B, n, m, p(xyzs[li]), p(uidx), p(kidx), p(f["idx"]), p(f["w"]), st), "3nn")
            else:
                chk(lib.pab_three_nn_weights(B, n, m, p(xyzs[li]), p(xyzs[li + 1]), p(f["idx"]), p(f["w"]), st), "3nn")

    def _launch_dense(self, xyz0, ws):
        """Feature path — fused SA modules + SA_Layer attention, FP modules, NetVLAD levels, gated fc head."""
        net = self.net
        B, N, _ = xyz0.shape
        lib, st, p, chk = L.lib(), L.stream_ptr(), L.ptr, L.check
        xyz, feat, c = xyz0, xyz0, 3
        for sa, lv in zip(self.sa, ws["levels"]):
            n, m, k = lv["n"], lv["m"], sa["k"]
            chk(lib.pab_sa_module_forward(B, n, m, k, k, c, p(xyz), p(feat), p(lv["cidx"]), p(lv["nbr"]), sa["layers"].arr,
                                          sa["layers"].n, p(lv["pooled"]), p(None), st), "sa")
            arr = sa["att"]["arr"]
            chk(lib.pab_sa_layer_forward_p(B, m, sa["att"]["C"], p(lv["pooled"]), arr, attention.C_ptr_offset(arr, 1),
                                           attention.C_ptr_offset(arr, 2), p(lv["feat"]), p(ws["scratch"]),
                                           self.attention_precision, st), "sa_layer")
            xyz, feat, c = lv["new_xyz"], lv["feat"], sa["layers"].c_out

        feats = [xyz0] + [lv["feat"] for lv in ws["levels"]]
        ns = [N] + [lv["m"] for lv in ws["levels"]]
        known_feat = feats[-1]
        for li in range(len(self.fp) - 1, -1, -1):
            f = ws["fp"][li]
            n, m = ns[li], ns[li + 1]
            skip = feats[li]
            chk(lib.pab_fp_module_forward(B, n, m, known_feat.shape[2], skip.shape[2], p(known_feat), p(skip), p(f["idx"]), p(f["w"]),
                                          self.fp[li].arr, self.fp[li].n, p(f["out"]), st), "fp")
            known_feat = f["out"]

        # pyramid order of the reference: f0 = level-3 features (64 points) ... f3 = level-0 features (4096 points)
        fp_out = [ws["fp"][i]["out"] for i in range(len(self.fp) - 1, -1, -1)]
        flat, off = ws["vlad"], 0
        for x_l, lvl in zip(fp_out, self.vlad):
            K, Cf = lvl["K"], lvl["C"]
            if K % 4 == 0:
                dst = C.c_void_p(flat.data_ptr() + 4 * off)       # element (c, k) of this level sits at off + c*K + k
                if Cf == 256:
                    chk(lib.pab_netvlad_forward_tc(B, x_l.shape[1], Cf, K, p(x_l), p(lvl["wc_hi"]),
                                                   p(lvl["wc_lo"] if self.vlad_precision == "f32" else None), p(lvl["shift"]),
                                                   p(lvl["w2"]), dst, flat.stride(0), K, p(ws["scratch"]), st), "vlad")
                else:
                    chk(lib.pab_netvlad_forward(B, x_l.shape[1], Cf, K, p(x_l), p(lvl["wc"]), p(lvl["shift"]), p(lvl["w2"]), dst,
                                                flat.stride(0), K, p(ws["scratch"]), st), "vlad")
            else:
                # a level with a cluster count that is not a multiple of four (K = 1 on 64 points): padded problem, real columns kept
                pad = ws["vlad_pad"]
                chk(lib.pab_netvlad_forward(B, x_l.shape[1], Cf, lvl["K4"], p(x_l), p(lvl["wc4"]), p(lvl["shift4"]), p(lvl["w24"]),
                                            p(pad), pad.stride(0), pad.stride(1), p(ws["scratch"]), st), "vlad")
                flat[:, off:off + Cf * K] = pad[:, :, :K].reshape(B, Cf * K)
            off += Cf * K
        gw, gs, gb = self.gate if self.gate is not None else (None, None, None)
        if lib.pab_gated_fc_tc_supported(self.flat, self.c_out):
            chk(lib.pab_gated_fc_forward_tc(B, self.flat, self.c_out, p(flat), p(self.fc_planes[0]), p(self.fc_planes[1]), p(self.fc_scale),
                                            p(self.fc_shift), p(gw), p(gs), p(gb), 1 if net.use_normalize else 0, p(ws["desc"]),
                                            p(ws["scratch"]), st), "head")
        else:
            chk(lib.pab_gated_fc_forward(B, self.flat, self.c_out, p(flat), p(self.fc_wt), p(self.fc_scale), p(self.fc_shift), p(gw), p(gs),
                                         p(gb), 1 if net.use_normalize else 0, p(ws["desc"]), p(ws["scratch"]), st), "head")
        return fp_out

    # ---- forward -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, return_feat=True):
        """x: (B,1,N,3) or (B,N,3) float32 CUDA -> (desc (B,256), fp_features [4 x (B,256,n,1)], center_idx_origin [4])."""
        L.require_cuda(x)
        xyz0 = (x.squeeze(1) if x.dim() == 4 else x).contiguous().float()
        B, N, _ = xyz0.shape
        ws = self._workspace(B, N)
        self._launch_geo(xyz0, ws)
        fp_out = self._launch_dense(xyz0, ws)
        out = ws["desc"].clone()
        if not return_feat:
            return out
        cidx = [lv["cidx"] for lv in ws["levels"]]
        origin = [cidx[0].clone()]
        for ci in cidx[1:]:
            origin.append(torch.gather(origin[-1], -1, ci.long()))                    # pptnet.py:109-118
        fp_features = [f.transpose(1, 2).unsqueeze(-1).clone() for f in fp_out]
        return out, fp_features, origin

    @torch.no_grad()
    def forward_stream(self, batches, out=None, ready_events=None, coalesce=0):
        """Throughput mode (as engine.FusedPatchAugNet.forward_stream): descriptors of a sequence of equally shaped batches, the
        geometry of batch i+1 (FPS is a serial chain on B of the 148 SMs) on a second stream under the dense kernels of batch i;
        two workspaces ping-pong, events order their reuse.  ``coalesce`` (clouds, 0 = off): consecutive batches are concatenated
        into launch sequences of up to that many clouds — bit-identical descriptors (every kernel's arithmetic depends on the cloud
        only), 28.8 k -> 32.7 k submaps/s (fp32 contract) and 31.2 k -> 35.4 k (bf16 mode) with 128 instead of 64 clouds per
        sequence.  ``ready_events[i]`` (optional): a CUDA event batch i's geometry waits for (its upload on a copy stream).
        Returns (len(batches)*B, c_out) on the device."""
        batches = list(batches)
        if not batches:
            return torch.empty(0, self.c_out, device=self.device)
        x0 = batches[0].squeeze(1) if batches[0].dim() == 4 else batches[0]
        B, N, _ = x0.shape
        if out is None:
            out = torch.empty(len(batches) * B, self.c_out, dtype=torch.float32, device=self.device)
        g = int(coalesce) // B if coalesce else 0
        cur = torch.cuda.current_stream()
        if g >= 2 and len(batches) >= g:
            n_groups = len(batches) // g
            merged, events = [], []
            for gi in range(n_groups):
                if ready_events is not None:
                    for ev in ready_events[gi * g:(gi + 1) * g]:
                        if ev is not None:
                            cur.wait_event(ev)
                merged.append(torch.cat([(x.squeeze(1) if x.dim() == 4 else x).float() for x in batches[gi * g:(gi + 1) * g]]))
                ev = torch.cuda.Event()
                ev.record(cur)
                events.append(ev)
            self.forward_stream(merged, out=out[:n_groups * g * B], ready_events=events)
            if len(batches) > n_groups * g:
                self.forward_stream(batches[n_groups * g:], out=out[n_groups * g * B:],
                                    ready_events=None if ready_events is None else ready_events[n_groups * g:])
            return out
        if getattr(self, "_streams", None) is None:
            self._streams = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        s_geo, s_dense = self._streams
        s_geo.wait_stream(cur)
        s_dense.wait_stream(cur)
        slots = [self._workspace(B, N, slot) for slot in (0, 1)]
        geo_done, dense_done = [None, None], [None, None]
        # as engine.FusedPatchAugNet.forward_stream: the persistent tensor-core kernels draw their tiles from a counter (a CTA whose
        # SM is held by an FPS CTA of the other stream starts late and takes fewer), the small-CTA SA0 kernel leaves registers free
        dyn = getattr(self, "stream_dynamic_tiles", True)
        if dyn:
            n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
            L.lib().pab_tune_tc_max_ctas(n_sm)
            L.lib().pab_tune_tensor_core(1 | 8)
        try:                                  # the tuning state is process-global: restore it whatever happens
            for i, x in enumerate(batches):
                L.require_cuda(x)
                xyz0 = (x.squeeze(1) if x.dim() == 4 else x).contiguous().float()
                slot = i & 1
                ws = slots[slot]
                with torch.cuda.stream(s_geo):
                    if ready_events is not None and ready_events[i] is not None:
                        s_geo.wait_event(ready_events[i])
                    if dense_done[slot] is not None:
                        s_geo.wait_event(dense_done[slot])             # workspace free again
                    self._launch_geo(xyz0, ws)
                    geo_done[slot] = torch.cuda.Event()
                    geo_done[slot].record()
                with torch.cuda.stream(s_dense):
                    s_dense.wait_event(geo_done[slot])
                    self._launch_dense(xyz0, ws)
                    out[i * B:(i + 1) * B].copy_(ws["desc"], non_blocking=True)
                    dense_done[slot] = torch.cuda.Event()
                    dense_done[slot].record()
                xyz0.record_stream(s_geo)
                xyz0.record_stream(s_dense)
            cur.wait_stream(s_dense)
            cur.wait_stream(s_geo)
        finally:
            if dyn:
                L.lib().pab_tune_tc_max_ctas(0)
                L.lib().pab_tune_tensor_core(1)
        return out

    __call__ = forward
