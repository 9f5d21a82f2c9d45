"""Layer schedules of the detector and the pose net over the C-ABI kernels (hn_b200.ops).

The nn.Modules in fcos_utils/ and a2j/ own the parameters (ordinary nn.Parameters / buffers with the
reference's state-dict keys).  The executors here derive a packed bf16 copy of the weights (rebuilt when a
parameter changes) and a set of statically shaped activation buffers per (batch, canvas) and then enqueue
the kernels layer by layer on the current CUDA stream.  No torch arithmetic runs on the hot path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .ops import Act, PhaseAct

BN_EPS = 1e-5      # FrozenBatchNorm2d / nn.BatchNorm2d default eps
GN_EPS = 1e-5
STEM_K_RGB = 256   # 8 kernel rows x 8 pixels x 4 channels (147 real taps), see hn_im2col_7x7s2
STEM_K_DEPTH = 64  # 8 kernel rows x 8 pixels (49 real taps)


def _version(t: torch.Tensor) -> int:
    # tensors created under torch.inference_mode() carry no version counter (reading it raises); they cannot be edited in
    # place outside inference mode either, so a constant is a correct fingerprint for them
    return 0 if t.is_inference() else t._version


def _sig(tensors: Sequence[torch.Tensor]):
    return tuple((t.data_ptr(), _version(t), str(t.device), tuple(t.shape)) for t in tensors)


class WeightsEpochMixin:
    """Change detection for the derived (packed bf16) weights and captured CUDA graphs.

    ``weights_epoch()`` changes whenever the module's parameters or buffers may have changed:
      * ``load_state_dict`` -- on the module itself or on any PARENT (``HandNet.load_state_dict``,
        ``A2JModelLightning.load_state_dict``): a load-state-dict post hook fires in nested loads too, where the
        children's ``load_state_dict`` override never runs;
      * ``.to()/.cuda()/.half()`` (``_apply``, reached through the parent's recursion as well);
      * in-place edits (``p.mul_()``, ``p.data.copy_()``, optimiser steps): the sum of the tensors' version counters
        is part of the epoch (~70 us for the ~650 tensors of both nets; fingerprinting pointers and shapes as well cost
        0.7 ms per step and is not needed: replacing a Parameter object goes through ``_apply`` or ``load_state_dict``).
    """

    _w_epoch = 0
    _w_tensors = None

    def _install_weight_hooks(self):
        # nn.Module.__init__ has run: hooks can be registered (called from the subclasses' __init__)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_weights())

    def invalidate_weights(self):
        self._w_epoch = self._w_epoch + 1
        self._w_tensors = None

    def weights_epoch(self):
        if self._w_tensors is None:
            self._w_tensors = list(self.parameters()) + list(self.buffers())
        return (self._w_epoch, sum(_version(t) for t in self._w_tensors))

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_weights()
        return out


def bn_affine(weight, bias, mean, var, eps=BN_EPS, conv_bias=None):
    """Eval-mode BatchNorm as y = x*scale + shift (torchvision/ops/misc.py:54-63), conv bias folded in."""
    scale = weight.float() * (var.float() + eps).rsqrt()
    shift = bias.float() - mean.float() * scale
    if conv_bias is not None:
        shift = shift + scale * conv_bias.float()
    return scale.contiguous(), shift.contiguous()


PLAN_SLOT = 0               # buffer set in use: steps that are in flight at the same time (GraphedHandNet(slot=...)) must not
                            # share activation buffers
PHASES = None               # set to a list to collect (name, CUDA event) marks of a step (tools/phase_timing.py)


def mark(name: str):
    """Record a timing event on the current stream (only while PHASES is a list)."""
    if PHASES is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PHASES.append((name, ev))


_side_streams: Dict[str, List["torch.cuda.Stream"]] = {}
SPLIT_K = True              # split-K (fp32 scratch + last-arriver epilogue) for the short, deep A2J layers
FUSE_LEVELS = True          # FCOS towers / output convolutions / GroupNorm: ONE launch over P3+P4+P5 per layer
                            # (hn_conv2d_bf16_levels) instead of one per level; False = round-1 schedule (six chains), kept for
                            # A/B timing and the equality test (outputs are bit-identical either way)
import os as _os
POSE_PRIORITY = _os.environ.get("HN_POSE_PRIO", "high")      # priority of the pose stream relative to the detect stream
STEM_WINDOW = _os.environ.get("HN_STEM_WINDOW", "1") != "0"    # detector stem: plain canvas + window descriptors (0: row-pair frame)
# CTAs per fused 256-wide tower convolution launch (HN_TOWER_CTAS; -1 = 80 % of the SMs, 0 = no cap).  These launches are not
# limited by the number of SMs (8 VGA frames, one B200: 148 / 140 / 124 / 108 / 98 / 84 CTAs -> towers phase 1126 / 1089 / 1083 /
# 1085 / 1114 / 1202 us): capped to ~80 % they are as fast, and the SMs they leave run the GroupNorm pass of the other tower
# and the pose stage of the previous step next to them.
TOWER_CTA_CAP = int(_os.environ.get("HN_TOWER_CTAS", "-1"))
DET_CTA_CAP = int(_os.environ.get("HN_DET_CTAS", "0"))       # CTAs per detector convolution launch (0 = all SMs): leaves SMs to the pose stage
POSE_CTA_CAP = int(_os.environ.get("HN_POSE_CTAS", "0"))     # CTAs per pose-net convolution launch (0 = all SMs)
POSE_PDL = _os.environ.get("HN_POSE_PDL", "1") != "0"        # programmatic dependent launch inside the pose stage
DET_PDL = _os.environ.get("HN_DET_PDL", "1") != "0"
PARALLEL_CHAINS = True      # independent layer chains (head towers, A2J towers) on forked streams / graph branches


def run_chains(chains, device):
    """Run independent callables on forked CUDA streams and join them on the current stream.  Under CUDA-graph
    capture this becomes a fork/join in the graph, so kernels of different chains fill each other's partial
    waves (most layers here have fewer tiles than a multiple of the 148 SMs)."""
    if not PARALLEL_CHAINS or len(chains) <= 1:
        for c in chains:
            c()
        return
    key = str(device)
    pool = _side_streams.setdefault(key, [])
    while len(pool) < len(chains) - 1:
        pool.append(torch.cuda.Stream(device=device))
    main = torch.cuda.current_stream(device)
    fork = torch.cuda.Event()
    fork.record(main)
    joins = []
    for i, c in enumerate(chains):
        if i == 0:
            continue
        st = pool[i - 1]
        st.wait_event(fork)
        with torch.cuda.stream(st):
            c()
            ev = torch.cuda.Event()
            ev.record(st)
            joins.append(ev)
    chains[0]()
    for ev in joins:
        main.wait_event(ev)


class ConvLayer:
    """Packed weights + epilogue vectors of one convolution."""

    __slots__ = ("w", "cout", "k", "stride", "dil", "scale", "shift")

    def __init__(self, weight, *, stride=1, dil=1, scale=None, shift=None):
        self.cout, _, self.k, _ = weight.shape
        self.w = ops.pack_conv_weight(weight)
        self.stride, self.dil = stride, dil
        self.scale = None if scale is None else scale.detach().float().contiguous()
        self.shift = None if shift is None else shift.detach().float().contiguous()

    def run(self, x, **kw):
        return ops.conv2d(x, self.w, cout=self.cout, ksize=self.k, stride=self.stride, dilation=self.dil,
                          scale=self.scale, shift=self.shift, **kw)


# =================================================================================================
# FCOS detector
# =================================================================================================
def resized_size(h: int, w: int, min_size: int, max_size: int) -> Tuple[int, int]:
    """torchvision transform.py:25-72 (eager branch): scale in python doubles, floor(in * scale)."""
    scale = min(min_size / min(h, w), max_size / max(h, w))
    return int(math.floor(float(h) * scale)), int(math.floor(float(w) * scale))


class FCOSWeights:
    def __init__(self, model):
        sd = {k: v for k, v in model.state_dict().items()}
        g = lambda k: sd[k]
        bn = lambda p: bn_affine(g(p + ".weight"), g(p + ".bias"), g(p + ".running_mean"), g(p + ".running_var"))
        b = "backbone.body."
        s, sh = bn(b + "bn1")
        self.stem_w = ops.pack_stem_weight(g(b + "conv1.weight"), STEM_K_RGB, order="window" if STEM_WINDOW else "pairs")
        self.stem_scale, self.stem_shift = s, sh
        self.blocks: List[List[Dict[str, ConvLayer]]] = []
        for li, nblocks in enumerate((3, 4, 6, 3), start=1):
            layer = []
            for bi in range(nblocks):
                p = f"{b}layer{li}.{bi}."
                stride = 2 if (li > 1 and bi == 0) else 1
                blk = {}
                s1, b1 = bn(p + "bn1")
                blk["conv1"] = ConvLayer(g(p + "conv1.weight"), stride=stride, scale=s1, shift=b1)
                s2, b2 = bn(p + "bn2")
                blk["conv2"] = ConvLayer(g(p + "conv2.weight"), scale=s2, shift=b2)
                if (p + "downsample.0.weight") in sd:
                    sd_, bd_ = bn(p + "downsample.1")
                    blk["down"] = ConvLayer(g(p + "downsample.0.weight"), stride=stride, scale=sd_, shift=bd_)
                layer.append(blk)
            self.blocks.append(layer)
        f = "backbone.fpn."
        self.inner = [ConvLayer(g(f"{f}inner_blocks.{i}.0.weight"), shift=g(f"{f}inner_blocks.{i}.0.bias")) for i in range(3)]
        self.outer = [ConvLayer(g(f"{f}layer_blocks.{i}.0.weight"), shift=g(f"{f}layer_blocks.{i}.0.bias")) for i in range(3)]
        self.towers = {}
        for name, hp in (("cls", "head.classification_head."), ("reg", "head.regression_head.")):
            layers = []
            for i in range(4):
                conv = ConvLayer(g(f"{hp}conv.{3 * i}.weight"), shift=g(f"{hp}conv.{3 * i}.bias"))
                layers.append((conv, g(f"{hp}conv.{3 * i + 1}.weight").float().contiguous(),
                               g(f"{hp}conv.{3 * i + 1}.bias").float().contiguous()))
            self.towers[name] = layers
        # fused output convs: [cls | hand_lr | contact | dydx(ReLU)] and [bbox_reg(ReLU) | ctrness]
        cp, rp = "head.classification_head.", "head.regression_head."
        names = [cp + "cls_logits", cp + "hand_lr_layer"]
        if model.ext:
            names += [cp + "hand_contact_state_layer", cp + "hand_dydx_layer"]
        self.cls_out = ConvLayer(torch.cat([g(n + ".weight") for n in names]), shift=torch.cat([g(n + ".bias") for n in names]))
        nc = g(cp + "cls_logits.weight").shape[0]
        self.num_classes = nc
        self.cls_cols = {"cls": (0, nc), "lr": (nc, nc + 2)}
        self.cls_relu = False
        if model.ext:
            self.cls_cols.update({"contact": (nc + 2, nc + 7), "dxdy": (nc + 7, nc + 10)})
            self.cls_relu = (nc + 7, nc + 10)                       # fcos_utils/fcos.py:301
        self.cls_ld = 8 if self.cls_out.cout <= 8 else 16
        self.reg_out = ConvLayer(torch.cat([g(rp + "bbox_reg.weight"), g(rp + "bbox_ctrness.weight")]),
                                 shift=torch.cat([g(rp + "bbox_reg.bias"), g(rp + "bbox_ctrness.bias")]))
        self.reg_relu = (0, 4)                                      # fcos_utils/fcos.py:379
        self.reg_ld = 8


class FCOSPlan:
    """Statically shaped buffers for one (batch, canvas) configuration."""

    def __init__(self, wts: FCOSWeights, batch: int, canvas_hw: Tuple[int, int], device, anchor_sizes):
        hc, wc = canvas_hw
        assert hc % 32 == 0 and wc % 32 == 0
        B = batch
        self.batch, self.canvas_hw = B, canvas_hw
        # the stem convolution gathers its 7x7 patches straight from the canvas (no im2col buffer): a plain row-major canvas
        # read through window descriptors, or (HN_STEM_WINDOW=0) the zero-framed row-pair canvas of rounds 1-2
        self.frame = ops.StemCanvas(B, (hc, wc), device) if STEM_WINDOW else ops.StemFrame(B, (hc, wc), device)
        self.canvas = self.frame.t            # (device handle for the chain launcher; read pixels with frame.canvas())
        h1, w1 = hc // 2, wc // 2
        self.stem = Act(B, h1, w1, 64, 0, device)
        sizes = [(hc // 4, wc // 4, 64), (hc // 8, wc // 8, 128), (hc // 16, wc // 16, 256), (hc // 32, wc // 32, 512)]
        self.stage = []
        for (h, w, c) in sizes:
            self.stage.append([Act(B, h, w, c, 1, device) for _ in range(3)])
        # phase-split copies of the outputs of layer1..3 feed the stride-2 convs of layer2..4
        self.phase = [PhaseAct(B, h, w, c, 1, device) for (h, w, c) in sizes[:3]]
        self.lvl_hw = [(h, w) for (h, w, _) in sizes[1:]]
        self.inner = [Act(B, h, w, 256, 1, device) for (h, w) in self.lvl_hw]
        self.p = [Act(B, h, w, 256, 1, device) for (h, w) in self.lvl_hw]
        self.tower = {t: [[Act(B, h, w, 256, 1, device) for _ in range(2)] for (h, w) in self.lvl_hw] for t in ("cls", "reg")}
        self.levels = ops.Levels(self.lvl_hw, canvas_hw, anchor_sizes)
        L = self.levels.locs
        self.locs = L
        # fused head outputs as fp32 channel PLANES [B][channel][locs] (hn_conv_desc.out_kind 2): coalesced stores in the
        # convolution epilogue, and the decode kernel streams just the planes it needs (cls, ctr) without padding bytes
        # (plane pitch rounded up to 32 locations: every plane starts 128-byte aligned, 16-byte loads over 4 locations)
        self.loc_pitch = (L + 31) // 32 * 32
        self.cls_buf = torch.zeros((B, wts.cls_ld, self.loc_pitch), dtype=torch.float32, device=device)
        self.reg_buf = torch.zeros((B, wts.reg_ld, self.loc_pitch), dtype=torch.float32, device=device)
        self.gn_stats = torch.zeros((2, 4, 3, B, 32, 2), dtype=torch.int64, device=device)    # fixed-point sums (ops.GN_FIX_SCALE)
        self.sel_ws = torch.empty(int(ops._lib.load().hn_fcos_select_workspace_bytes(B, L)), dtype=torch.uint8, device=device)
        self.nms_ws = ops.nms_workspace(B, L, device)


class FCOSExecutor:
    """Eval-mode FCOS.forward (fcos_utils/fcos.py:675-767) on the GPU."""

    def __init__(self, model):
        self.model = model
        self._wsig = None
        self.wts: Optional[FCOSWeights] = None
        self.plans: Dict[Tuple, FCOSPlan] = {}

    def weights(self) -> FCOSWeights:
        sig = self.model.weights_epoch()
        if sig != self._wsig or self.wts is None:
            self.wts = FCOSWeights(self.model)
            self._wsig = sig
            self.plans.clear()
        return self.wts

    def plan(self, batch, canvas_hw, device) -> FCOSPlan:
        key = (batch, canvas_hw, str(device), PLAN_SLOT)
        if key not in self.plans:
            self.plans[key] = FCOSPlan(self.wts, batch, canvas_hw, device, self.model.anchor_sizes)
        return self.plans[key]

    # ------------------------------------------------------------------------------------------
    def backbone_heads(self, pl: FCOSPlan):
        """canvas -> fused fp32 head buffers."""
        w = self.wts
        B = pl.batch
        hc, wc = pl.canvas_hw
        h1, w1 = hc // 2, wc // 2
        mark("preprocess")
        ops.conv2d(pl.frame, w.stem_w, cout=64, ksize=1, scale=w.stem_scale, shift=w.stem_shift, relu=True, out=pl.stem,
                   algo_k=147)
        ops.maxpool3x3s2(pl.stem.t, pl.stage[0][0])
        feats = []
        mark("stem+maxpool")
        for li in range(4):
            bufs = pl.stage[li]
            # the layer as a list of convolutions over buffer indices (three rotating buffers per stage)
            prog = []
            xi = 0 if li == 0 else None          # layer1 starts from the pooled stem in bufs[0]; later layers from the phase copy
            for bi, blk in enumerate(w.blocks[li]):
                last = bi == len(w.blocks[li]) - 1
                ph = last and li < 3
                if "down" in blk:
                    f0, f1 = (0, 1) if xi is None else [i for i in range(3) if i != xi][:2]
                    prog.append(dict(conv=blk["conv1"], src="phase", relu=True, out=f0, res=None, out_phase=False))
                    prog.append(dict(conv=blk["down"], src="phase", relu=False, out=f1, res=None, out_phase=False))
                    # the identity buffer is read and overwritten by the same threads of the epilogue
                    prog.append(dict(conv=blk["conv2"], src=f0, relu=True, out=f1, res=f1, out_phase=ph))
                    xi = f1
                else:
                    f0, f1 = [i for i in range(3) if i != xi][:2]
                    prog.append(dict(conv=blk["conv1"], src=xi, relu=True, out=f0, res=None, out_phase=False))
                    prog.append(dict(conv=blk["conv2"], src=f0, relu=True, out=f1, res=xi, out_phase=ph))
                    xi = f1

            def emit(ins, bset, phase_in, phase_out):
                src = phase_in if ins["src"] == "phase" else bset[ins["src"]]
                ins["conv"].run(src, relu=ins["relu"], out=bset[ins["out"]],
                                res=None if ins["res"] is None else bset[ins["res"]], res_mode=1 if ins["res"] is not None else 0,
                                out_phase=phase_out if ins["out_phase"] else None)

            phase_in = pl.phase[li - 1] if li > 0 else None
            phase_out = pl.phase[li] if li < 3 else None
            # (Running a badly quantised layer -- layer3 at 8 VGA frames: 228 tiles = two rounds at 77 % -- as two half-batch
            # chains on two streams, so that one chain's next convolution takes the SMs the other leaves idle, was measured:
            # layer3 388 -> 382 us, step unchanged within noise; not kept.)
            for ins in prog:
                emit(ins, bufs, phase_in, phase_out)
            x = bufs[xi]
            if li >= 1:
                feats.append(x)
            mark(f"layer{li + 1}")
        # FPN top-down (torchvision/ops/feature_pyramid_network.py:172-204)
        for i in (2, 1, 0):
            if i == 2:
                w.inner[i].run(feats[i], out=pl.inner[i])
            else:
                w.inner[i].run(feats[i], res=pl.inner[i + 1], res_mode=2, out=pl.inner[i])
        run_chains([(lambda i=i: w.outer[i].run(pl.inner[i], out=pl.p[i])) for i in range(3)], pl.canvas.device)
        mark("fpn")
        # heads (fcos_utils/fcos.py:267-329, 373-395)
        pl.gn_stats.zero_()

        def tower_chain(ti, t, lvl):
            def run():
                x = pl.p[lvl]
                for i, (conv, gamma, beta) in enumerate(w.towers[t]):
                    o = pl.tower[t][lvl][i & 1]
                    st = pl.gn_stats[ti, i, lvl]
                    conv.run(x, out=o, gn_stats=st, gn_groups=32)
                    ops.groupnorm_relu(o, st, 32, gamma, beta, GN_EPS)
                    x = o
                if t == "cls":
                    w.cls_out.run(x, relu=w.cls_relu, out_f32=pl.cls_buf, out_rows_per_image=pl.loc_pitch,
                                  out_row_offset=pl.levels.starts[lvl], out_planar=True)
                else:
                    w.reg_out.run(x, relu=w.reg_relu, out_f32=pl.reg_buf, out_rows_per_image=pl.loc_pitch,
                                  out_row_offset=pl.levels.starts[lvl], out_planar=True)
            return run

        tower_cap = TOWER_CTA_CAP if TOWER_CTA_CAP >= 0 else (ops.device_info()[0] * 4 // 5) & ~1

        def tower_chain_levels(ti, t):
            # the reference applies the SAME tower modules to every level (fcos_utils/fcos.py:278-289, 378-380): one launch per
            # layer over the concatenated tile list of P3+P4+P5 (1169 tiles for 8 VGA frames = 8 full waves of 147 CTAs)
            # instead of three launches of which the P4 / P5 ones fill the 148 SMs to 77 % / 41 %
            def run():
                xs = list(pl.p)
                nl = len(xs)
                for i, (conv, gamma, beta) in enumerate(w.towers[t]):
                    outs = [pl.tower[t][lvl][i & 1] for lvl in range(nl)]
                    sts = [pl.gn_stats[ti, i, lvl] for lvl in range(nl)]
                    if tower_cap:
                        ops.conv_cta_cap(tower_cap)
                    ops.conv2d_levels(xs, conv.w, cout=conv.cout, ksize=conv.k, shift=conv.shift, outs=outs, gn_stats=sts,
                                      gn_groups=32)
                    if tower_cap:
                        ops.conv_cta_cap(DET_CTA_CAP)
                    ops.groupnorm_relu_levels(outs, sts, 32, gamma, beta, GN_EPS)
                    xs = outs
                oc, relu, buf = (w.cls_out, w.cls_relu, pl.cls_buf) if t == "cls" else (w.reg_out, w.reg_relu, pl.reg_buf)
                ops.conv2d_levels(xs, oc.w, cout=oc.cout, ksize=oc.k, shift=oc.shift, relu=relu, out_f32=buf,
                                  out_rows_per_image=pl.loc_pitch, out_row_offsets=pl.levels.starts[:nl], out_planar=True)
            return run

        if FUSE_LEVELS and len(pl.p) <= 3:
            run_chains([tower_chain_levels(ti, t) for ti, t in enumerate(("cls", "reg"))], pl.canvas.device)
        else:
            # six independent chains: (cls, reg) x (P3, P4, P5); the big P3 chains first
            run_chains([tower_chain(ti, t, lvl) for lvl in range(3) for ti, t in enumerate(("cls", "reg"))],
                       pl.canvas.device)
        mark("towers")
        return pl.cls_buf, pl.reg_buf

    def head_views(self, pl: FCOSPlan):
        """The reference's head tensors [B, locs, k] as (strided) views of the channel-planar buffers."""
        c = self.wts.cls_cols
        cv = lambda a, b: pl.cls_buf[:, a:b, :pl.locs].permute(0, 2, 1)
        v = {"cls_logits": cv(*c["cls"]), "hand_lr": cv(*c["lr"]),
             "bbox_regression": pl.reg_buf[:, 0:4, :pl.locs].permute(0, 2, 1), "bbox_ctrness": pl.reg_buf[:, 4:5, :pl.locs].permute(0, 2, 1)}
        if "contact" in c:
            v["hand_contact_state"] = cv(*c["contact"])
            v["hand_dxdy_relu"] = cv(*c["dxdy"])
        return v

    def postprocess(self, pl: FCOSPlan, ratios_h, ratios_w):
        """fcos_utils/fcos.py:572-669 on the fused head buffers -> dense per-image detections on the device."""
        m = self.model
        v = self.head_views(pl)
        cand = ops.fcos_decode_select(v["cls_logits"], v["bbox_ctrness"], v["bbox_regression"], self.wts.num_classes,
                                      pl.levels, m.score_cut, ws=pl.sel_ws)
        keep, keep_count = ops.nms_batched(cand["box"], cand["score"], cand["label"], cand["count"], m.nms_iou,
                                           m.nms_coord_trick_numel, ws=pl.nms_ws)
        out = ops.fcos_gather(keep, keep_count, cand, v["hand_lr"], pl.levels, ratios_h, ratios_w,
                              contact=v.get("hand_contact_state"), dxdy=v.get("hand_dxdy_relu"))
        out["keep_count"] = keep_count
        out["cand_count"] = cand["count"]
        out["cand"] = cand
        out["keep"] = keep
        return out

    def forward_device(self, images: Sequence[torch.Tensor]):
        """Whole detector; returns dense device tensors (capacity = #locations) + keep_count, no host sync."""
        m = self.model
        self.weights()
        dev = images[0].device
        orig = [(int(im.shape[-2]), int(im.shape[-1])) for im in images]
        sizes = [resized_size(h, w, m.min_size, m.max_size) for h, w in orig]
        hc = int(math.ceil(max(s[0] for s in sizes) / 32.0) * 32)
        wc = int(math.ceil(max(s[1] for s in sizes) / 32.0) * 32)
        pl = self.plan(len(images), (hc, wc), dev)
        if isinstance(pl.frame, ops.StemCanvas):
            ops.preprocess(images, sizes, (hc, wc), m.image_mean, m.image_std, canvas=pl.frame.t)
        else:
            ops.preprocess(images, sizes, (hc, wc), m.image_mean, m.image_std, frame=pl.frame.t)
        self.backbone_heads(pl)
        # resize_boxes ratios are float32 tensor divisions in the reference (fcos_utils/fcos.py:771-776)
        f32 = torch.float32
        rh = [float(torch.tensor(o[0], dtype=f32) / torch.tensor(s[0], dtype=f32)) for o, s in zip(orig, sizes)]
        rw = [float(torch.tensor(o[1], dtype=f32) / torch.tensor(s[1], dtype=f32)) for o, s in zip(orig, sizes)]
        out = self.postprocess(pl, rh, rw)
        mark("postprocess")
        out["plan"] = pl
        return out


# =================================================================================================
# A2J pose net
# =================================================================================================
class A2JWeights:
    def __init__(self, model):
        sd = model.state_dict()
        g = lambda k: sd[k]

        def bn(p, conv_bias=None):
            return bn_affine(g(p + ".weight"), g(p + ".bias"), g(p + ".running_mean"), g(p + ".running_var"),
                             conv_bias=conv_bias)
        b = "Backbone.model."
        w0 = g(b + "conv1.weight").float()
        self.channel_in = model.Backbone.channel_in
        if self.channel_in == 1:
            # x.expand(n,3,h,w) of the depth channel == convolving with the weights summed over Cin (a2j/a2j.py:197-199)
            w0 = w0.sum(1, keepdim=True)
            self.stem_k = STEM_K_DEPTH
        else:
            # RGBD variant (a2j/a2j.py:191-192): a 4-channel 7x7 stem, run as the direct stem over a framed NHWC4 canvas
            assert self.channel_in == 4 and w0.shape[1] == 4, "A2J stem: 1 (depth) or 4 (RGBD) input channels"
            self.stem_k = STEM_K_RGB
        self.stem_w = ops.pack_stem_weight(w0, self.stem_k)
        self.stem_scale, self.stem_shift = bn(b + "bn1")
        self.blocks = []
        for li, nblocks in enumerate((3, 4, 6, 3), start=1):
            layer = []
            for bi in range(nblocks):
                p = f"{b}layer{li}.{bi}."
                stride = 2 if (li in (2, 3) and bi == 0) else 1       # stride on the 3x3 (a2j/resnet.py:68)
                dil = 2 if (li == 4 and bi > 0) else 1                 # a2j/resnet.py:112,142,145
                blk = {"stride": stride, "dil": dil}
                s, sh = bn(p + "bn1")
                blk["conv1"] = ConvLayer(g(p + "conv1.weight"), scale=s, shift=sh)
                s, sh = bn(p + "bn2")
                blk["conv2"] = ConvLayer(g(p + "conv2.weight"), stride=stride, dil=dil, scale=s, shift=sh)
                s, sh = bn(p + "bn3")
                blk["conv3"] = ConvLayer(g(p + "conv3.weight"), scale=s, shift=sh)
                if (p + "downsample.0.weight") in sd:
                    s, sh = bn(p + "downsample.1")
                    blk["down"] = ConvLayer(g(p + "downsample.0.weight"), stride=stride, scale=s, shift=sh)
                layer.append(blk)
            self.blocks.append(layer)
        self.towers = {}
        for name in ("classificationModel", "regressionModel", "DepthRegressionModel"):
            if not hasattr(model, name):
                continue
            layers = []
            for i in range(1, 5):
                s, sh = bn(f"{name}.bn{i}", conv_bias=g(f"{name}.conv{i}.bias"))
                layers.append(ConvLayer(g(f"{name}.conv{i}.weight"), scale=s, shift=sh))
            out = ConvLayer(g(f"{name}.output.weight"), shift=g(f"{name}.output.bias"))
            self.towers[name] = (layers, out)
        self.anchors = g("post_process.all_anchors").float().contiguous()


class A2JPlan:
    def __init__(self, wts: A2JWeights, n: int, hw: Tuple[int, int], device, num_joints: int):
        h, w = hw
        self.n = n
        h1, w1 = (h + 1) // 2, (w + 1) // 2
        if wts.channel_in == 4:
            assert h % 2 == 0 and w % 2 == 0, "RGBD crops must have even sizes"
            self.frame = ops.StemFrame(n, (h, w), device)
        else:
            self.stem_a = torch.empty((n * h1 * w1, wts.stem_k), dtype=torch.bfloat16, device=device)
        self.stem = Act(n, h1, w1, 64, 0, device)
        h2, w2 = (h1 + 1) // 2, (w1 + 1) // 2
        self.pool = Act(n, h2, w2, 64, 1, device)
        A = lambda hh, ww, c, halo=1: Act(n, hh, ww, c, halo, device)
        self.bufs = {}
        self.hw = [(h2, w2), ((h2 + 1) // 2, (w2 + 1) // 2)]
        self.hw.append(((self.hw[1][0] + 1) // 2, (self.hw[1][1] + 1) // 2))
        self.hw.append(self.hw[2])
        self.A = A
        self.device = device
        hf, wf = self.hw[3]
        self.feat_hw = (hf, wf)
        na = hf * wf * 16
        self.cls = torch.zeros((n, na, num_joints), dtype=torch.float32, device=device)
        self.reg = torch.zeros((n, na, num_joints, 2), dtype=torch.float32, device=device)
        self.dep = torch.zeros((n, na, num_joints), dtype=torch.float32, device=device)
        self.agg_ws = torch.empty(int(ops._lib.load().hn_a2j_workspace_bytes(n, num_joints)), dtype=torch.uint8, device=device)
        self.cache: Dict[str, object] = {}

    def act(self, key, hh, ww, c, halo=1):
        if key not in self.cache:
            self.cache[key] = Act(self.n, hh, ww, c, halo, self.device)
        return self.cache[key]

    def phase(self, key, hh, ww, c):
        if key not in self.cache:
            self.cache[key] = PhaseAct(self.n, hh, ww, c, 1, self.device)
        return self.cache[key]

    def splitk(self, key, x, conv: "ConvLayer"):
        """Split-K scratch (one fp32 slice of the padded output per K split + zeroed tile counters) for a convolution
        reading `x`, or None when the layer has enough tiles on its own.  Exclusive per convolution: independent
        convolutions run concurrently."""
        if not SPLIT_K:
            return None
        hh, ww = (x.h2, x.w2) if isinstance(x, PhaseAct) else (x.h, x.w)
        rows = self.n * (hh + 2 * x.halo) * (ww + 2 * x.halo)
        m_tiles = (rows + 127) // 128
        cout_pad = conv.w.shape[0]
        if m_tiles * (cout_pad // 64) * 2 > 148:
            return None
        k = "sk_" + key
        if k not in self.cache:
            tiles = m_tiles * max(1, cout_pad // 64)
            slices = max(2, min(16, 148 // tiles))          # the library uses as many K splits as slices fit
            self.cache[k] = (torch.empty(slices * m_tiles * 128 * cout_pad, dtype=torch.float32, device=self.device),
                             torch.zeros(m_tiles * (cout_pad // 16), dtype=torch.int32, device=self.device))
        return self.cache[k]


class A2JExecutor:
    """A2JModel.forward with gt=None (a2j/a2j.py:243-250) on the GPU."""

    def __init__(self, model):
        self.model = model
        self._wsig = None
        self.wts: Optional[A2JWeights] = None
        self.plans: Dict[Tuple, A2JPlan] = {}

    def weights(self) -> A2JWeights:
        sig = self.model.weights_epoch()
        if sig != self._wsig or self.wts is None:
            self.wts = A2JWeights(self.model)
            self._wsig = sig
            self.plans.clear()
        return self.wts

    def heads_device(self, x: torch.Tensor):
        """x: fp32 [n, 1, H, W] on the device -> (cls, reg, dep) fp32 head tensors in the reference layout."""
        w = self.weights()
        n, _, h, wd = x.shape
        key = (n, h, wd, str(x.device), PLAN_SLOT)
        if key not in self.plans:
            self.plans[key] = A2JPlan(w, n, (h, wd), x.device, self.model.num_joints)
        pl = self.plans[key]
        if w.channel_in == 4:
            # x[:, 0:4] (a2j/a2j.py:196) -> framed bf16 NHWC4 -> direct stem
            ops.pack_nhwc4_frame(x.contiguous(), (0, 1, 2, 3), pl.frame)
            ops.conv2d(pl.frame, w.stem_w, cout=64, ksize=1, scale=w.stem_scale, shift=w.stem_shift, relu=True, out=pl.stem,
                       algo_k=196)
        else:
            depth = x[:, 0].contiguous() if x.shape[1] != 1 else x.reshape(n, h, wd)
            _, h1, w1 = ops.im2col_7x7s2(depth, w.stem_k, out=pl.stem_a)
            a = Act(n, h1, w1, w.stem_k, 0, x.device, t=pl.stem_a.view(n, h1, w1, w.stem_k))
            ops.conv2d(a, w.stem_w, cout=64, ksize=1, scale=w.stem_scale, shift=w.stem_shift, relu=True, out=pl.stem,
                       algo_k=147)       # 3 identical input channels in the reference: 7*7*3 MACs per output
        cur = ops.maxpool3x3s2(pl.stem.t, pl.pool)
        self._schedule(pl, w, cur, n)
        return pl.cls, pl.reg, pl.dep, pl

    def _schedule(self, pl: A2JPlan, w: A2JWeights, cur, n):
        """Emit the convolutions of layer1..layer4 and the towers in dependency order."""
        c4 = None
        for li in range(4):
            planes = 64 << li
            for bi, blk in enumerate(w.blocks[li]):
                hin, win = cur.h, cur.w
                stride, dil = blk["stride"], blk["dil"]
                hout, wout = ((hin + 1) // 2, (win + 1) // 2) if stride == 2 else (hin, win)
                tag = f"l{li}b{bi}"
                # group 1: conv1 (1x1) and, if present, the downsample conv -- both read the block input
                if stride == 2:
                    # conv1 runs at the input resolution and also writes the phase-split copy the stride-2 3x3 reads;
                    # the stride-2 downsample reads the phase-split copy of the block input
                    t1 = pl.act(tag + "t1", hin, win, planes)
                    t1p = pl.phase(tag + "t1p", hin, win, planes)
                    blk["conv1"].run(cur, relu=True, out=t1, out_phase=t1p, splitk=pl.splitk(tag + "c1", cur, blk["conv1"]))
                    idn = pl.act(tag + "id", hout, wout, planes * 4)
                    blk["down"].run(pl.cache[f"l{li - 1}out_phase"], out=idn,
                                    splitk=pl.splitk(tag + "dn", pl.cache[f"l{li - 1}out_phase"], blk["down"]))
                    src2 = t1p
                else:
                    t1 = pl.act(tag + "t1", hin, win, planes, dil)       # a dilated 3x3 needs a halo of 2 on its input
                    blk["conv1"].run(cur, relu=True, out=t1, splitk=pl.splitk(tag + "c1", cur, blk["conv1"]))
                    if "down" in blk:
                        idn = pl.act(tag + "id", hout, wout, planes * 4)
                        blk["down"].run(cur, out=idn, splitk=pl.splitk(tag + "dn", cur, blk["down"]))
                    else:
                        idn = cur
                    src2 = t1
                t2 = pl.act(tag + "t2", hout, wout, planes)
                blk["conv2"].run(src2, relu=True, out=t2, splitk=pl.splitk(tag + "c2", src2, blk["conv2"]))
                last = bi == len(w.blocks[li]) - 1
                out = pl.act(tag + "out", hout, wout, planes * 4)
                out_phase = pl.phase(f"l{li}out_phase", hout, wout, planes * 4) if (last and li in (0, 1)) else None
                blk["conv3"].run(t2, relu=True, res=idn, res_mode=1, out=out, out_phase=out_phase,
                                 splitk=pl.splitk(tag + "c3", t2, blk["conv3"]))
                cur = out
            if li == 2:
                c4 = cur
        c5 = cur
        hf, wf = pl.feat_hw
        towers = (("regressionModel", c5, pl.reg), ("DepthRegressionModel", c5, pl.dep), ("classificationModel", c4, pl.cls))
        if ops.PROFILE is None and PARALLEL_CHAINS:
            # eager / per-layer launches: the three towers are independent chains -> forked streams (graph branches)
            def chain(name, src, dst):
                def run():
                    x = src
                    for i in range(4):
                        o = pl.act(f"{name}{i}", hf, wf, 256)
                        w.towers[name][0][i].run(x, relu=True, out=o, splitk=pl.splitk(f"{name}{i}", x, w.towers[name][0][i]))
                        x = o
                    w.towers[name][1].run(x, out_f32=dst.view(n, hf * wf, -1), out_transpose_hw=True,
                                          splitk=pl.splitk(name + "out", x, w.towers[name][1]))
                return run
            run_chains([chain(*t) for t in towers], pl.device)
            return
        t = {name: src for name, src, _ in towers}
        for i in range(4):
            for name, _, _ in towers:
                o = pl.act(f"{name}{i}", hf, wf, 256)
                w.towers[name][0][i].run(t[name], relu=True, out=o,
                                         splitk=pl.splitk(f"{name}{i}", t[name], w.towers[name][0][i]))
                t[name] = o
        for name, _, dst in towers:
            # permute(0,3,2,1) + view of the reference == w-major rows (a2j/a2j.py:85-89)
            w.towers[name][1].run(t[name], out_f32=dst.view(n, hf * wf, -1), out_transpose_hw=True,
                                  splitk=pl.splitk(name + "out", t[name], w.towers[name][1]))

    def forward_device(self, x: torch.Tensor) -> torch.Tensor:
        cls, reg, dep, pl = self.heads_device(x)
        mark("a2j convs")
        out = ops.a2j_aggregate(cls, reg, dep, self.wts.anchors, ws=pl.agg_ws)
        mark("a2j aggregate")
        return out


# =================================================================================================
# whole-step executor: static buffers + CUDA graph
# =================================================================================================
RECORD_WIDTH = 21 * 3 + 4 + 1      # joints, crop box, has_hand  


def weights_token(net) -> int:
    """Epochs of the two models (see WeightsEpochMixin): changes when weights are reloaded or moved."""
    pose = net.a2j.a2j if hasattr(net.a2j, "a2j") else net.a2j
    return (net.detector.weights_epoch(), pose.weights_epoch())


def pack_records(joints: torch.Tensor, crops: torch.Tensor, has_hand: torch.Tensor) -> torch.Tensor:
    """Fixed-size per-frame result record [B, 68] fp32: 63 joint coords, 4 crop ints (exact in fp32), hit flag."""
    b = joints.shape[0]
    return torch.cat((joints.reshape(b, -1).float(), crops.float(), has_hand.float().reshape(b, 1)), dim=1)


def unpack_records(rec: torch.Tensor):
    joints = rec[:, :63].reshape(-1, 21, 3)
    crops = rec[:, 63:67].round().to(torch.int64)
    has = rec[:, 67] > 0.5
    return joints, crops, has


class GraphedHandNet:
    """The step over fixed (batch, H, W) input buffers as a TWO-STAGE PIPELINE of CUDA graphs:

      detect stage  (stream `det`):   preprocess -> FCOS -> post-process -> hand select + crop     (~110 launches)
      pose stage    (stream `pose`):  A2J pose net -> anchor aggregation -> per-frame records       (~75 launches)

    The pose net of 8 crops is a chain of ~52 dependent, latency-bound launches that keeps well under a third of the SMs
    busy (16 % of a step for 4 % of its FLOPs); it depends on the detector of the SAME step only.  `submit()` therefore
    enqueues the detect stage of step i on `det` and the pose stage on the higher-priority `pose` stream, so that the
    pose stage of step i runs under the detect stage of step i+1.  The hand-off (crops, hit flags, 176x176 depth crops;
    1 MB) is copied into a pose-private buffer at the start of the pose stage, and the detect stage of step i+1 waits for
    that copy, so the stages share no buffer.  `result(ticket)` waits for one step's records on the host.  Results are
    bit-identical to running the two stages back to back on one stream (same kernels, same buffers, same order per stage).
    """

    RING = 4            # steps in flight at most (pinned result buffers)

    def __init__(self, net, batch: int, h: int, w: int, depth_c: int = 1, use_graph: bool = True, slot: int = 0):
        self.net = net
        self.slot = slot            # buffer set: give concurrent executors (one per stream) different slots
        dev = next(net.parameters()).device
        self.dev = dev
        self.batch = batch
        self.hands = int(getattr(net, "max_hands", 1))     # hand slots per frame: the pose stage runs batch * hands crops
        self.rgb = torch.zeros((batch, 3, h, w), dtype=torch.float32, device=dev)
        self.depth = torch.zeros((batch, depth_c, h, w), dtype=torch.float32, device=dev)
        self.images = list(self.rgb.unbind(0))
        self.use_graph = use_graph
        self.g_det = self.g_pose = None
        self.token = None
        self.out = None
        self.rec = None
        self.launches_per_step = 0
        lo, hi = torch.cuda.Stream.priority_range()            # (lowest, highest) = (0, -N)
        pr = {"high": (lo, hi), "low": (hi, lo), "same": (lo, lo)}[POSE_PRIORITY]
        self.det_stream = torch.cuda.Stream(device=dev, priority=pr[0])
        self.pose_stream = torch.cuda.Stream(device=dev, priority=pr[1])
        # hand-off blob [crops int64 B*4 | has_hand int32 B (padded to 8 B) | depth crops fp32] twice: written by the detect
        # stage / read by the pose stage
        from handnet_pipeline.handnet_pipeline import CROP_SIZE
        self.crop = CROP_SIZE
        slots = batch * self.hands
        self._off_has = slots * 4 * 8
        self._off_depth = self._off_has + ((slots * 4 + 7) // 8) * 8
        nbytes = self._off_depth + slots * depth_c * CROP_SIZE * CROP_SIZE * 4
        self.hand_d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.hand_p = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.depth_c = depth_c
        self.ring_host = [torch.empty((slots, RECORD_WIDTH), dtype=torch.float32).pin_memory() for _ in range(self.RING)]
        self.ring_done = [None] * self.RING
        self.ring_out = [None] * self.RING
        self.rec_host = self.ring_host[0]
        self.d2h_bytes = self.rec_host.numel() * 4
        self.ev_copy = None         # hand-off of the latest submitted step has been copied (detect stage may overwrite it)
        self.copy_stream = None     # host -> device uploads (created on first use)
        self._stage = {}            # staging sets for host inputs, by input kind
        self.n_submitted = 0
        self.n_collected = 0

    # ---------------------------------------------------------------------------------------------- buffers
    def _views(self, blob):
        b, c, s = self.batch * self.hands, self.depth_c, self.crop
        crops = blob[:self._off_has].view(torch.int64).view(b, 4)
        has = blob[self._off_has:self._off_has + b * 4].view(torch.int32)
        depth = blob[self._off_depth:].view(torch.float32).view(b, c, s, s)
        return crops, has, depth

    def load_inputs(self, rgb, depth: torch.Tensor):
        """Copy a batch (rgb: [B,3,H,W] tensor or list of [3,H,W]; depth [B,C,H,W]) into the static input buffers, on the
        detect stream, ordered after the caller's current stream.  HOST tensors (pinned memory for a truly asynchronous
        copy) are uploaded on a separate copy stream into one of two staging sets, so that the PCIe transfer of step i+1
        overlaps the kernels of step i; the detect stream then takes them with a device-to-device copy."""
        first = rgb[0] if isinstance(rgb, (list, tuple)) else rgb
        if first.device.type == "cpu":
            return self._load_staged(rgb, depth, u8=False)
        cur = torch.cuda.current_stream(self.dev)
        self.det_stream.wait_stream(cur)
        with torch.cuda.stream(self.det_stream):
            if isinstance(rgb, (list, tuple)):
                torch._foreach_copy_(self.images, list(rgb))
            else:
                self.rgb.copy_(rgb, non_blocking=True)
            self.depth.copy_(depth, non_blocking=True)
        # the sources are read on the detect stream, possibly long after this call has returned: tell the caching allocator,
        # or a caller's temporary (`x.cuda()`) could be recycled for the next batch's upload before it has been copied
        for t in (list(rgb) if isinstance(rgb, (list, tuple)) else [rgb]) + [depth]:
            if t.is_cuda:
                t.record_stream(self.det_stream)

    def load_frames_u8(self, bgr_u8: torch.Tensor, depth_u16: torch.Tensor):
        """Camera frames (uint8 BGR [B,H,W,3], uint16 / int16 millimetres [B,H,W]; host or device) -> the static fp32 input
        buffers through hn_ingest_frames (ros_demo.py:227-238, 266-267); host frames are staged like load_inputs."""
        if bgr_u8.device.type == "cpu":
            return self._load_staged(bgr_u8, depth_u16, u8=True)
        self.det_stream.wait_stream(torch.cuda.current_stream(self.dev))
        bgr_u8, depth_u16 = bgr_u8.contiguous(), depth_u16.contiguous()
        with torch.cuda.stream(self.det_stream):
            ops.ingest_frames(bgr_u8, depth_u16, rgb_out=self.rgb, depth_out=self.depth)
        bgr_u8.record_stream(self.det_stream)
        depth_u16.record_stream(self.det_stream)

    def _load_staged(self, a, b, u8: bool):
        key = "u8" if u8 else "f32"
        if key not in self._stage:
            if u8:
                mk = lambda: (torch.empty((self.batch,) + tuple(self.rgb.shape[2:]) + (3,), dtype=torch.uint8, device=self.dev),
                              torch.empty((self.batch,) + tuple(self.rgb.shape[2:]), dtype=b.dtype, device=self.dev))
            else:
                mk = lambda: (torch.empty_like(self.rgb), torch.empty_like(self.depth))
            self._stage[key] = {"bufs": [mk(), mk()], "free": [None, None], "n": 0}
        st = self._stage[key]
        k = st["n"] & 1
        st["n"] += 1
        sa, sb = st["bufs"][k]
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.dev)
        cs = self.copy_stream
        if st["free"][k] is not None:
            cs.wait_event(st["free"][k])                    # the detect stream has consumed this staging set
        with torch.cuda.stream(cs):
            if isinstance(a, (list, tuple)):
                for dst, src in zip(sa.unbind(0), a):
                    dst.copy_(src, non_blocking=True)
            else:
                sa.copy_(a, non_blocking=True)
            sb.copy_(b.reshape(sb.shape), non_blocking=True)
            up = torch.cuda.Event()
            up.record(cs)
        with torch.cuda.stream(self.det_stream):
            self.det_stream.wait_event(up)
            if u8:
                ops.ingest_frames(sa, sb, rgb_out=self.rgb, depth_out=self.depth)
            else:
                self.rgb.copy_(sa, non_blocking=True)
                self.depth.copy_(sb, non_blocking=True)
            fr = torch.cuda.Event()
            fr.record(self.det_stream)
            st["free"][k] = fr

    # ---------------------------------------------------------------------------------------------- stages
    def _stage_detect(self):
        global PLAN_SLOT
        saved, PLAN_SLOT = PLAN_SLOT, self.slot
        try:
            ops.conv_cta_cap(DET_CTA_CAP)
            ops.conv_pdl(DET_PDL)
            det, crops, has, depth_batch = self.net.detect_crop_device(self.images, self.depth, out=self._views(self.hand_d))
        finally:
            ops.conv_cta_cap(0)
            ops.conv_pdl(True)
            PLAN_SLOT = saved
        self.det = det
        return det

    def _stage_pose(self):
        global PLAN_SLOT
        saved, PLAN_SLOT = PLAN_SLOT, self.slot
        crops, has, depth_batch = self._views(self.hand_p)
        try:
            ops.conv_cta_cap(POSE_CTA_CAP)
            ops.conv_pdl(POSE_PDL)
            joints = self.net.pose_device(depth_batch)
        finally:
            ops.conv_cta_cap(0)
            ops.conv_pdl(True)
            PLAN_SLOT = saved
        self.rec = pack_records(joints, crops, has)
        self.out = {"joints": joints, "has_hand": has, "crops": crops, "depth_batch": depth_batch, "det": self.det}
        return self.out

    def _eager(self):
        """Both stages back to back on the CURRENT stream (tools, conv_profile, capture warm-up)."""
        self._stage_detect()
        self.hand_p.copy_(self.hand_d)
        return self._stage_pose()

    def capture(self):
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        l0 = ops.launch_count()
        self.g_det = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_det):
            self._stage_detect()
        self.g_pose = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_pose):
            self._stage_pose()
        self.launches_per_step = ops.launch_count() - l0
        self.token = weights_token(self.net)
        self.det_stream.wait_stream(cur)
        self.pose_stream.wait_stream(cur)

    def invalidate_if_weights_changed(self):
        if self.g_det is not None and self.token != weights_token(self.net):
            self.drain()
            self.g_det = self.g_pose = None

    # ---------------------------------------------------------------------------------------------- pipeline
    def submit(self, want_outputs: bool = False, post=None) -> int:
        """Enqueue one step over the loaded inputs; returns its ticket.  Never waits for the device.  At most RING steps may
        be in flight (collect the oldest one with result() first).  `post(rec)` runs on the pose stream right after the
        step's kernels with the device records [B, 68] (hn_b200.parallel enqueues its all-gather there)."""
        if self.n_submitted - self.n_collected >= self.RING:
            raise RuntimeError(f"GraphedHandNet: {self.RING} steps in flight; call result() before submitting more")
        if self.use_graph and self.g_det is None:
            self.capture()
        ticket = self.n_submitted
        r = ticket % self.RING
        det, pose = self.det_stream, self.pose_stream
        with torch.cuda.stream(det):
            if self.ev_copy is not None:
                det.wait_event(self.ev_copy)              # the previous step's pose stage has taken its hand-off
            if self.use_graph:
                self.g_det.replay()
            else:
                l0 = ops.launch_count()
                self._stage_detect()
            ev_det = torch.cuda.Event()
            ev_det.record(det)
        with torch.cuda.stream(pose):
            pose.wait_event(ev_det)
            self.hand_p.copy_(self.hand_d, non_blocking=True)
            self.ev_copy = torch.cuda.Event()
            self.ev_copy.record(pose)
            if self.use_graph:
                self.g_pose.replay()
            else:
                self._stage_pose()
                self.launches_per_step = ops.launch_count() - l0
            self.ring_host[r].copy_(self.rec, non_blocking=True)
            if post is not None:
                post(self.rec)
            if want_outputs:                               # private copies: the next step's hand-off overwrites hand_p
                self.ring_out[r] = (self.out["depth_batch"].clone(), self.out["crops"].clone())
            done = torch.cuda.Event()
            done.record(pose)
            self.ring_done[r] = done
        self.n_submitted += 1
        return ticket

    def result(self, ticket: int):
        """Wait (on the host) for step `ticket`; returns (rec_host [B, 68] pinned fp32, (depth_batch, crops) or None).
        Tickets are collected in submission order."""
        if ticket != self.n_collected or ticket >= self.n_submitted:
            raise RuntimeError(f"GraphedHandNet.result: ticket {ticket} is not the oldest step in flight ({self.n_collected})")
        r = ticket % self.RING
        self.ring_done[r].synchronize()
        self.n_collected += 1
        self.rec_host = self.ring_host[r]
        outs, self.ring_out[r] = self.ring_out[r], None
        return self.ring_host[r], outs

    def done_event(self, ticket: int):
        return self.ring_done[ticket % self.RING]

    def drain(self):
        while self.n_collected < self.n_submitted:
            self.result(self.n_collected)

    def run(self):
        """One step, both stages, finished before the call returns control of the buffers: the CURRENT stream waits for the
        pose stage (no host sync)."""
        self.drain()
        t = self.submit()
        torch.cuda.current_stream(self.dev).wait_event(self.done_event(t))
        self.n_collected += 1              # nothing to collect on the host: the caller reads device buffers
        return self.out

    def records(self) -> torch.Tensor:
        return self.rec

    def run_e2e(self, rgb_host: torch.Tensor, depth_host: torch.Tensor):
        """Host (pinned) frames in, host results out: H2D + step + D2H + one host wait."""
        self.drain()
        self.load_inputs(rgb_host, depth_host)
        rec, _ = self.result(self.submit())
        return unpack_records(rec)

    def counts(self):
        d = self.out["det"]
        return {"cand": d["cand_count"].tolist(), "kept": d["keep_count"].tolist(),
                "hands": int(self.out["has_hand"].sum().item())}


def conv_profile(step: GraphedHandNet, repeats: int = 3):
    """Summed device time of the tcgen05 conv launches of one step (ms) and their number.

    The step is enqueued eagerly behind a long spin kernel so that the host finishes queueing before the GPU
    starts: the CUDA events around each conv launch (on the launch stream) then bracket back-to-back kernels
    and contain no host-induced gaps."""
    global PARALLEL_CHAINS
    best = None
    n = 0
    saved, PARALLEL_CHAINS = PARALLEL_CHAINS, False      # one stream: per-launch times must not overlap
    for _ in range(repeats):
        ops.PROFILE = []
        torch.cuda._sleep(150_000_000)
        step._eager()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        ms = sum(p[0].elapsed_time(p[1]) for p in prof)
        n = len(prof)
        step.last_conv_flops = sum(p[2] for p in prof)
        if best is None or ms < best:
            best = ms
            step.last_conv_table = [dict(p[3], ms=p[0].elapsed_time(p[1]), gflop=p[2] / 1e9) for p in prof]
    PARALLEL_CHAINS = saved
    return best, n


def conv_flops_per_step(step: GraphedHandNet) -> float:
    """Algorithmic FLOPs (2*MAC over interior output pixels, real K) of the conv launches of one step."""
    return float(step.last_conv_flops)
