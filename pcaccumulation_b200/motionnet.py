"""MotionNet: drop-in replacement of the reference's ``models.motionnet.MotionNet`` on B200.

Same constructor (``MotionNet(cfg)`` reading the reference's config keys), same ``state_dict`` key names
and shapes (195 tensors, so released checkpoints load through ``toolbox/utils.py:partial_load``), same
``forward(input_dict) -> results`` contract (``models/motionnet.py:137-262``).  The ``torch.nn`` layers
declared here only HOLD the parameters; every stage of ``forward`` runs as hand-written sm_100a CUDA behind
the C ABI in ``include/pcab200.h`` (``libpcab200.so``).  There is no CPU / PyTorch fallback: without the
shared library or a CUDA device ``forward`` raises.

Internal layout: all BEV tensors are NHWC float32 ``[B*T, Ny, Nx, C]``; points are additionally indexed
in pillar-sorted order so per-pillar reductions are contiguous ranges.
"""
import ctypes
import math

import numpy as np

import torch
import torch.nn as nn

from . import _lib as L
from ._lib import D, F, I, P, Z, call, host_floats, scratch, size, stream

MIN_POINTS = 15  # models/motionnet.py:11
N_KPTS = 1024


# ------------------------------------------------------------------------------------------------
# parameter containers (names/shapes == reference; no compute happens in these modules)
# ------------------------------------------------------------------------------------------------
class _ResBlockParams(nn.Module):  # models/pillar_encoder.py:13-44
    def __init__(self, size_in, size_out):
        super().__init__()
        self.fc_0 = nn.Linear(size_in, min(size_in, size_out))
        self.fc_1 = nn.Linear(min(size_in, size_out), size_out)
        self.shortcut = nn.Linear(size_in, size_out, bias=False)
        nn.init.zeros_(self.fc_1.weight)


class _PillarEncoderParams(nn.Module):  # models/pillar_encoder.py:59-95
    def __init__(self, cfg):
        super().__init__()
        nf = cfg["num_filters"]
        self.fc_pos = nn.Linear(cfg["num_input_features"], 2 * nf)
        self.fc_c = nn.Linear(nf, nf)
        self.blocks = nn.ModuleList([_ResBlockParams(2 * nf, nf) for _ in range(cfg["depth"])])


class _Down(nn.Module):  # models/unet.py:45-62
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)


class _Up(nn.Module):  # models/unet.py:74-97
    def __init__(self, cin, cout):
        super().__init__()
        self.upconv = nn.ConvTranspose2d(cin, cout, 2, stride=2)
        self.conv1 = nn.Conv2d(2 * cout, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)


class _UNetParams(nn.Module):  # models/unet.py:116-220
    def __init__(self, in_channels=3, depth=5, start_filts=64, **kwargs):
        super().__init__()
        downs, ups = [], []
        outs = in_channels
        for i in range(depth):
            ins = in_channels if i == 0 else outs
            outs = start_filts * (2 ** i)
            downs.append(_Down(ins, outs))
        for i in range(depth - 1):
            ins = outs
            outs = ins // 2
            ups.append(_Up(ins, outs))
        self.down_convs = nn.ModuleList(downs)
        self.up_convs = nn.ModuleList(ups)
        self.conv_final = nn.Conv2d(outs, in_channels, 3, padding=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)


class _SegHead2DParams(nn.Module):  # models/unet.py:259-277
    def __init__(self, cin, cout):
        super().__init__()
        mid = max(cin, cout)
        self.seg_head = nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1), nn.BatchNorm2d(mid), nn.ReLU(),
                                      nn.Conv2d(mid, cout, 3, padding=1))


class _SegHead1DParams(nn.Module):  # models/unet.py:235-256
    def __init__(self, cin, cout):
        super().__init__()
        mid = max(cin, cout)
        self.seg_head = nn.Sequential(nn.Linear(cin, mid), nn.BatchNorm1d(mid), nn.ReLU(), nn.Linear(mid, cout))


class _EgoHeadParams(nn.Module):  # models/egomotion.py:35-42
    def __init__(self):
        super().__init__()
        self.beta = nn.Parameter(torch.tensor(-5.0))
        self.alpha = nn.Parameter(torch.tensor(-5.0))


class _STPNParams(nn.Module):  # models/stpn.py:7-59
    def __init__(self, feat=32):
        super().__init__()
        widths = [32, 64, 128, 128, 256]
        self.init_conv = nn.Sequential(*[m for _ in range(4) for m in (nn.Conv3d(feat if _ == 0 else widths[0], widths[0], 3, padding=1), nn.ReLU())])
        downs, ins = [], feat
        for w in widths:
            w = max(64, w)
            downs.append(_Down(ins, w))
            ins = w
        ups, ins = [], widths[-1]
        for w in widths[-2::-1]:
            w = max(64, w)
            ups.append(_Up(ins, w))
            ins = w
        self.down_convs = nn.ModuleList(downs)
        self.up_convs = nn.ModuleList(ups)
        self.positional_encoding = nn.Sequential(nn.Linear(3, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU())
        self.final_proj = nn.Sequential(nn.Linear(128, 128), nn.ReLU())
        self.mos_seg = _SegHead1DParams(128, 2)
        self.offset_head = _SegHead1DParams(128, 2)


class _TPointNetParams(nn.Module):  # models/tpointnet.py:171-205
    def __init__(self):
        super().__init__()
        def mlp(a, b, c, d):
            return nn.Sequential(nn.Linear(a, b), nn.ReLU(), nn.Linear(b, c), nn.ReLU(), nn.Linear(c, d))
        self.geo_embed = mlp(32, 32, 64, 128)
        self.motion_embed = mlp(64, 64, 128, 128)
        self.pos_embed = mlp(4, 32, 64, 128)
        self.regressor = nn.Sequential(nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Linear(256, 128),
                                       nn.BatchNorm1d(128), nn.ReLU(), nn.Linear(128, 7))


class _AlignNetParams(nn.Module):  # models/alignnet.py:44-51
    def __init__(self):
        super().__init__()
        self.alignment = _TPointNetParams()


# ------------------------------------------------------------------------------------------------
# weight packing helpers (kernel-side layouts are documented next to each kernel)
# ------------------------------------------------------------------------------------------------
def _t(w):  # Linear weight [out,in] -> [in][out]
    return w.detach().float().t().contiguous().reshape(-1)


def _v(b):
    return b.detach().float().contiguous().reshape(-1)


def _bn_affine(bn):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def _pack_conv3x3(weight, splits):
    """[Cout, Cin, 3, 3] -> concatenated per-source blocks [9][C_s][Cout]."""
    w = weight.detach().float().permute(2, 3, 1, 0).reshape(9, weight.shape[1], weight.shape[0])
    blocks, c0 = [], 0
    for c in splits:
        blocks.append(w[:, c0:c0 + c, :].contiguous().reshape(-1))
        c0 += c
    return torch.cat(blocks).contiguous()


def _pack_conv3d(weight):
    """[Cout, Cin, 3, 3, 3] -> three blocks (kt = 0,1,2) of [9][Cin][Cout]."""
    return torch.cat([_pack_conv3x3(weight[:, :, kt], [weight.shape[1]]) for kt in range(3)]).contiguous()


def _pack_convT(weight):
    """[Cin, Cout, 2, 2] -> [4 = dy*2+dx][Cin][Cout]."""
    return weight.detach().float().permute(2, 3, 0, 1).contiguous().reshape(-1)


class _ConvLayer:
    """Packed weights of one 3x3 convolution (optionally with a BatchNorm(eval) epilogue)."""

    def __init__(self, conv, splits=None, bn=None, temporal=False):
        w = conv.weight
        self.cout = w.shape[0]
        self.temporal = temporal
        if temporal:
            self.splits = [w.shape[1]] * 3
            self.pack = _pack_conv3d(w)
        else:
            self.splits = splits or [w.shape[1]]
            self.pack = _pack_conv3x3(w, self.splits)
        self.bias = _v(conv.bias)
        self.bn = _bn_affine(bn) if bn is not None else (None, None)
        self.weight = w  # for the tensor-core pack (built lazily)
        self.tc_pack = None
        self.tc_pack16 = None
        self.p16_pack = None  # K-dense fp16-pair pack of the P16 kernel (tc_pack.pack_conv_p16)
        self.p16_pack3d = None  # Conv3d with fused temporal taps (tc_pack.pack_conv3d_fused_p16)
        self.tc_scale16 = None  # power-of-two scale of the fp16-pair weight pack (tc_pack.f16_weight_scale)
        self.tc_ok = {}  # (H, W, fmt) -> does the tensor-core kernel take this layer at that map size


class _MergedConv:
    """Two 3x3 convolutions over the SAME input run as one layer (output channels concatenated): the first convolutions of
    the two SegHead2D heads both read ``bev_feats``; merged, the input plane is staged once and the MMA N is 192 + 96."""

    def __init__(self, a, b):
        import types

        conv = types.SimpleNamespace(weight=torch.cat((a.weight.detach(), b.weight.detach()), 0), bias=torch.cat((a.bias, b.bias)))
        self.layer = _ConvLayer(conv)
        self.layer.bn = (torch.cat((a.bn[0], b.bn[0])).contiguous(), torch.cat((a.bn[1], b.bn[1])).contiguous())
        self.c_first = a.cout


class MotionNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        pe_cfg, unet_cfg = cfg["pillar_encoder"], cfg["unet"]
        assert pe_cfg["depth"] == 3 and pe_cfg["num_filters"] == 32 and pe_cfg["num_input_features"] == 9, \
            "the fused pillar encoder is specialised for depth 3 / 32 filters / 9 features (reference default)"
        assert unet_cfg["in_channels"] == 32 and cfg["pose_estimation"]["feats_dim"] == 64
        assert cfg["pose_estimation"]["n_kpts"] == N_KPTS and cfg["pose_estimation"]["add_slack"]
        self.pillar_encoder = _PillarEncoderParams(pe_cfg)
        self.unet = _UNetParams(**unet_cfg)
        self.semseg_head = _SegHead2DParams(unet_cfg["in_channels"], 2)
        self.ego_feats_head = _SegHead2DParams(unet_cfg["in_channels"], cfg["pose_estimation"]["feats_dim"])
        self.ego_motion_head = _EgoHeadParams()
        self.resolution = cfg["voxel_generator"]["voxel_size"]
        self.pc_range = cfg["voxel_generator"]["range"]
        self.motionhead = _STPNParams(cfg["stpn"]["feat_dim"])
        self.mode = cfg["misc"]["mode"]
        self.reconstructor = _AlignNetParams()
        self.n_sweeps = cfg["voxel_generator"]["n_sweeps"]
        self._packed = None
        self._packed_key = None
        self._plist = None
        self.use_tensor_cores = True
        # operand format of the tensor-core convolutions: "f16" = fp16 pairs (kind::f16, activations saturate at +-65504),
        # "tf32" = 3xTF32 (kind::tf32, full FP32 range, ~1.3x slower); both keep ~22 significant bits per operand
        self.conv_operands = "f16"
        # With fp16-pair operands the BEV activations themselves live in the pair-packed P16 format (csrc/pair16.cuh): every
        # producer writes (h, l) pairs, the convolutions read them straight into the MMA operand.  Values beyond +-65504 would
        # saturate: the producers count such events and forward() then repeats the scene with conv_operands = "tf32".
        self.packed_activations = True
        self.fuse_conv3d = True  # Conv3d 3x3x3 with the temporal taps in the MMA N dimension (P16 path)
        self.pfn_tensor_cores = True  # pillar encoder on tcgen05 (fp16-pair operands) when use_tensor_cores
        self.merge_heads = True  # first convolutions of semseg_head / ego_feats_head as one 96-channel layer (P16 path)
        self._sat = None  # device counter of saturated P16 outputs
        self._p16_ok = {}
        self.stages = {}  # stage-boundary tensors of the last forward (for stage-wise parity tests)
        self.keep_stages = False
        # stage-wise parity protocol: tensors placed here replace the computed value for the stages
        # DOWNSTREAM of it; keys: 'fb_est_map' [B,T,1,Ny,Nx], 'ego_motion_est' [B,T,4,4], 'mos_est' [N,2],
        # 'offset_est' [N,2], 'inst_labels_est' [N], 'transformed_points' [N,3] (consumed by clustering / TubeNet only)
        self.inject = {}
        self.stage_marks = None  # when a list: (name, cuda event) at stage boundaries (profiling aid)
        self.rng = None  # torch.Generator for the keypoint permutations (None = the global CPU generator, as upstream)
        self.conv_events = None  # when a list: (start_event, end_event, flops, path) per conv launch (bench roofline)
        # The two convolution stacks (backbone UNet + FG/BG trunk; Conv3d + STPN UNet) are fixed launch sequences over fixed
        # shapes: after one eager pass they are captured into CUDA graphs and replayed (2 graph launches instead of ~50 kernel
        # launches from Python per scene).  Their inputs are persistent buffers; their outputs are rewritten by the next forward.
        self.use_graphs = True
        self._graphs = {}
        self.flop_acc = None  # when a dict: algorithmic FLOPs of the tensor-core launches per stack name (bench roofline)
        self._flop_key = None

    # ------------------------------------------------------------------------------------------
    def _pack_key(self):
        # (storage, in-place version) of every parameter / buffer; the tensor list is cached (walking the module tree costs
        # ~1 ms per forward) and rebuilt whenever the module is moved / cast (``_apply`` replaces the tensors)
        if self._plist is None:
            self._plist = list(self.parameters()) + list(self.buffers())
        return tuple((p.data_ptr(), p._version) for p in self._plist)

    def _apply(self, fn, *args, **kwargs):
        self._plist = None
        self._static_weights = False
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._static_weights = False
        return super().load_state_dict(*args, **kwargs)

    def freeze_weights(self, frozen=True):
        """Serving mode: promise that the parameters stay as they are, so a forward reuses the packed kernel operands without
        comparing the (storage, version) of all 195 tensors first (~0.15 ms of interpreter time per scene, held under the GIL
        that the scenes in flight share).  ``load_state_dict`` / ``.to()`` / ``.half()`` lift the promise; in-place edits of a
        parameter do NOT -- call ``freeze_weights(False)`` before making them."""
        self._static_weights = bool(frozen) and self._packed is not None
        return self

    def _weights(self):
        if getattr(self, "_static_weights", False) and self._packed is not None:
            return self._packed
        key = self._pack_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        W = {}
        pe = self.pillar_encoder
        parts = [_t(pe.fc_pos.weight), _v(pe.fc_pos.bias)]
        for blk in pe.blocks:
            parts += [_t(blk.fc_0.weight), _v(blk.fc_0.bias), _t(blk.fc_1.weight), _v(blk.fc_1.bias), _t(blk.shortcut.weight)]
        parts += [_t(pe.fc_c.weight), _v(pe.fc_c.bias)]
        W["pfn"] = torch.cat(parts).contiguous()
        assert W["pfn"].numel() == L.lib().pcab_pfn_pack_size()
        from .tc_pack import pack_pfn_tc
        blob, bias_tc, inv = pack_pfn_tc(pe)
        W["pfn_tc"] = (blob, bias_tc, host_floats(inv))

        def unet_layers(net, prefix, final):
            for i, d in enumerate(net.down_convs):
                W[f"{prefix}d{i}c1"] = _ConvLayer(d.conv1)
                W[f"{prefix}d{i}c2"] = _ConvLayer(d.conv2)
            for i, u in enumerate(net.up_convs):
                co = u.upconv.weight.shape[1]
                W[f"{prefix}u{i}up"] = (_pack_convT(u.upconv.weight), _v(u.upconv.bias), u.upconv.weight.shape[0], co,
                                        {"weight": u.upconv.weight})
                W[f"{prefix}u{i}c1"] = _ConvLayer(u.conv1, splits=[co, u.conv1.weight.shape[1] - co])
                W[f"{prefix}u{i}c2"] = _ConvLayer(u.conv2)
            if final:
                W[f"{prefix}final"] = _ConvLayer(net.conv_final)

        unet_layers(self.unet, "unet.", True)
        unet_layers(self.motionhead, "stpn.", False)
        sh = self.semseg_head.seg_head
        W["sem0"] = _ConvLayer(sh[0], bn=sh[1])
        W["sem3"] = (sh[3].weight.detach().float().permute(2, 3, 1, 0).contiguous().reshape(-1), _v(sh[3].bias))
        eh = self.ego_feats_head.seg_head
        W["ego0"] = _ConvLayer(eh[0], bn=eh[1])
        W["ego3"] = _ConvLayer(eh[3])
        W["semego0"] = _MergedConv(W["sem0"], W["ego0"])
        for j, i in enumerate((0, 2, 4, 6)):
            W[f"stpn.c3d{j}"] = _ConvLayer(self.motionhead.init_conv[i], temporal=True)
        mh = self.motionhead
        ms, os_ = mh.mos_seg.seg_head, mh.offset_head.seg_head
        s_m, t_m = _bn_affine(ms[1])
        s_o, t_o = _bn_affine(os_[1])
        pad2 = torch.zeros(2, device=s_m.device)
        W["stpn_head"] = torch.cat([
            _t(mh.positional_encoding[0].weight), _v(mh.positional_encoding[0].bias),
            _t(mh.positional_encoding[2].weight), _v(mh.positional_encoding[2].bias),
            _t(mh.final_proj[0].weight), _v(mh.final_proj[0].bias),
            _t(ms[0].weight), _v(ms[0].bias), s_m, t_m, _t(ms[3].weight), _v(ms[3].bias), pad2,
            _t(os_[0].weight), _v(os_[0].bias), s_o, t_o, _t(os_[3].weight), _v(os_[3].bias), pad2]).contiguous()
        assert W["stpn_head"].numel() == L.lib().pcab_stpn_head_pack_size()
        from .tc_pack import pack_stpn_head_tc
        W["stpn_head_tc1"], W["stpn_head_tc"] = pack_stpn_head_tc(mh)
        W["stpn_head_host"] = W["stpn_head"].cpu().contiguous()  # small vectors go into the kernel parameter block
        al = self.reconstructor.alignment

        def mlp_pack(seq):
            return torch.cat([_t(seq[0].weight), _v(seq[0].bias), _t(seq[2].weight), _v(seq[2].bias),
                              _t(seq[4].weight), _v(seq[4].bias)]).contiguous()

        W["tpn_motion"], W["tpn_geo"], W["tpn_pos"] = mlp_pack(al.motion_embed), mlp_pack(al.geo_embed), mlp_pack(al.pos_embed)
        from .tc_pack import pack_embed_tc
        W["tpn_motion_tc"], W["tpn_geo_tc"] = pack_embed_tc(al.motion_embed), pack_embed_tc(al.geo_embed)
        W["tpn_pos_tc"] = pack_embed_tc(al.pos_embed, first=1)  # layer 0 (4 -> 32) stays on the CUDA cores
        r = al.regressor
        s0, t0 = _bn_affine(r[1])
        s1, t1 = _bn_affine(r[4])
        W["tpn_reg"] = torch.cat([_t(r[0].weight), _v(r[0].bias), s0, t0, _t(r[3].weight), _v(r[3].bias), s1, t1,
                                  _t(r[6].weight), _v(r[6].bias)]).contiguous()
        W["alpha"] = self.ego_motion_head.alpha.detach().float().reshape(1).contiguous()
        W["beta"] = self.ego_motion_head.beta.detach().float().reshape(1).contiguous()
        self._packed, self._packed_key = W, key
        return W

    # ------------------------------------------------------------------------------------------
    # activation format of the BEV tensors
    # ------------------------------------------------------------------------------------------
    def _fmt(self, B, T, Ny, Nx):
        """1 = P16 (pair-packed fp16, csrc/pair16.cuh) when the tensor-core path with fp16-pair operands is selected and every
        convolution of both stacks has a tile plan at this grid size; 0 = float32."""
        if not (self.use_tensor_cores and self.conv_operands == "f16" and self.packed_activations):
            return 0
        key = (Ny, Nx)
        ok = self._p16_ok.get(key)
        if ok is None:
            lib = L.lib()
            ok = Ny % 16 == 0 and Nx % 16 == 0
            h, w = Ny, Nx
            for _ in range(5):  # the five resolutions of both UNets; channel counts never matter for the plan's existence
                ok = ok and bool(lib.pcab_conv3x3_p16_supported(I(1), I(32), I(0), I(0), I(32), I(h), I(w)))
                h, w = h // 2, w // 2
            self._p16_ok[key] = ok
        return 1 if ok else 0

    def _sat_counter(self, dev):
        if self._sat is None or self._sat.device != dev:
            self._sat = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._sat

    # ------------------------------------------------------------------------------------------
    # CUDA-graph replay of fixed launch sequences
    # ------------------------------------------------------------------------------------------
    def _graph_input(self, key, shape, dev):
        """Persistent input buffer of the graphed stack ``key`` (the producer writes into it every forward)."""
        slot = self._graphs.get(key)
        if slot is None:
            slot = self._graphs[key] = {"in": torch.empty(*shape, device=dev, dtype=torch.float32), "graph": None, "out": None}
        return slot["in"]

    def _run_stack(self, key, fn, capture=False):
        """outputs = fn(static_input): replay of the captured graph when there is one for the current weights, eager
        otherwise.  Graphs are only captured by ``warmup()`` (stream capture while other host threads drive the same GPU
        is fragile), never implicitly."""
        slot = self._graphs[key]
        if slot.get("key") is not self._packed_key:  # first use, or a new weight pack drops the capture
            if slot.get("graph") is not None:
                import warnings

                warnings.warn("pcaccumulation_b200: weights changed after warmup(); the captured CUDA graphs are dropped and "
                              "the convolution stacks run kernel by kernel until warmup() is called again")
            slot.update(graph=None, out=None, key=self._packed_key)
        if self.flop_acc is not None:  # accounting pass: eager, FLOPs of this stack's launches summed under its name
            self._flop_key = key[0]
            out = fn(slot["in"])
            self._flop_key = None
            return out
        if not (self.use_graphs and self.conv_events is None):
            return fn(slot["in"])
        if slot["graph"] is not None:
            slot["graph"].replay()
            return slot["out"]
        if not capture:
            return fn(slot["in"])
        fn(slot["in"])  # eager pass first: lazy weight packs, kernel attributes, allocator warm-up
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            with L.pinned_stream():  # the capture runs on a side stream
                out = fn(slot["in"])
        slot["graph"], slot["out"] = g, out
        return out

    def _backbone_stack(self, W, B, T, Ny, Nx, fmt):
        def backbone(x):
            feats = self._unet(W, "unet.", x, B * T, Ny, Nx, self.cfg["unet"]["depth"], True, fmt)
            if fmt and self.merge_heads:  # [sem0 (32) | ego0 (64)] channels in one tensor
                return feats, self._conv(W["semego0"].layer, [feats], B * T, Ny, Nx, True, fmt=fmt)
            return feats, self._conv(W["sem0"], [feats], B * T, Ny, Nx, True, fmt=fmt)
        return backbone

    def _stpn_stack(self, W, B, T, Ny, Nx, dev, fmt):
        def stpn_stack(x):
            for j in range(4):
                x = self._conv(W[f"stpn.c3d{j}"], [x], B * T, Ny, Nx, True, T=T, fmt=fmt)
            xm = torch.empty(B, Ny, Nx, 32, device=dev)
            call("pcab_temporal_max", P(x), P(xm), I(B), I(T), I(Ny), I(Nx), I(32), I(fmt), stream())
            del x
            return self._unet(W, "stpn.", xm, B, Ny, Nx, 5, False, fmt)
        return stpn_stack

    def time_conv_stacks(self, reps=20):
        """CUDA-event time (ms per replay) of the captured graphs of the two convolution stacks, replayed back to back on the
        current stream over whatever their static inputs hold: the tensor-core convolutions with no host in the loop."""
        out = {}
        for key, slot in self._graphs.items():
            if slot.get("graph") is None:
                continue
            slot["graph"].replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                slot["graph"].replay()
            e1.record()
            e1.synchronize()
            out[key[0]] = e0.elapsed_time(e1) / reps
        return out

    @torch.no_grad()
    def warmup(self, batch_size=1):
        """Capture the CUDA graphs of the two convolution stacks for ``batch_size`` scenes per forward (grid and sweep
        count come from the config).  Call it from the thread that owns the model while no other thread is using the GPU
        (``SceneRunner`` / ``ScenePipeline`` do); without it every forward launches the stacks kernel by kernel."""
        if not self.use_graphs:
            return
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise L.PcabError("warmup() needs the model on a CUDA device")
        vg = self.cfg["voxel_generator"]
        Nx = int(round((vg["range"][3] - vg["range"][0]) / vg["voxel_size"][0]))
        Ny = int(round((vg["range"][4] - vg["range"][1]) / vg["voxel_size"][1]))
        B, T = int(batch_size), self.n_sweeps
        fmt = self._fmt(B, T, Ny, Nx)
        tc = (self.use_tensor_cores, self.conv_operands, fmt, self.merge_heads, self.fuse_conv3d)
        with L.pinned_stream():
            W = self._weights()
            self._sat_counter(dev)
            for name, fn in (("backbone", self._backbone_stack(W, B, T, Ny, Nx, fmt)), ("stpn", self._stpn_stack(W, B, T, Ny, Nx, dev, fmt))):
                key = (name, B, T, Ny, Nx, tc)
                self._graph_input(key, (B * T, Ny, Nx, 32), dev).zero_()
                self._run_stack(key, fn, capture=True)
        torch.cuda.current_stream().synchronize()
        self.freeze_weights()  # warmup() is the serving-mode entry: weights are static until load_state_dict / .to()

    # ------------------------------------------------------------------------------------------
    # convolution dispatch
    # ------------------------------------------------------------------------------------------
    def _conv(self, layer, srcs, n_img, H, W_, relu, out=None, T=1, fmt=0, src0_cstride=0, src0_off=0):
        dev = srcs[0].device
        if out is None:
            out = torch.empty(n_img, H, W_, layer.cout, device=dev, dtype=torch.float32)
        s = list(srcs) + [None] * (3 - len(srcs))
        c = list(layer.splits) + [0] * (3 - len(layer.splits))
        if layer.temporal:
            s = [srcs[0]] * 3
        scale, shift = layer.bn
        ev = self.conv_events
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        tail = (P(layer.bias), P(scale), P(shift), I(int(relu)), P(out), I(n_img), I(H), I(W_), I(layer.cout))
        Tl = I(T if layer.temporal else 1)
        tc_ok = False
        if self.use_tensor_cores and not fmt:
            tc_ok = layer.tc_ok.get((H, W_))
            if tc_ok is None:
                tc_ok = layer.tc_ok[(H, W_)] = bool(L.lib().pcab_conv3x3_tc_supported(
                    I(len(layer.splits)), I(c[0]), I(c[1]), I(c[2]), I(layer.cout), I(H), I(W_)))
        if (fmt or (tc_ok and self.conv_operands == "f16")) and layer.tc_scale16 is None:
            from .tc_pack import f16_weight_scale
            layer.tc_scale16 = f16_weight_scale(layer.weight)
        if fmt and layer.p16_pack is None and not (layer.temporal and self.fuse_conv3d and layer.cout == 32 and layer.splits[0] == 32):
            from .tc_pack import pack_conv_p16
            layer.p16_pack = pack_conv_p16(layer, layer.tc_scale16)
        if not fmt and tc_ok and self.conv_operands == "f16" and layer.tc_pack16 is None:
            from .tc_pack import pack_conv_tc_f16
            layer.tc_pack16 = pack_conv_tc_f16(layer, layer.tc_scale16)
        if fmt and layer.temporal and self.fuse_conv3d and layer.cout == 32 and layer.splits[0] == 32:
            # Conv3d with the temporal taps fused into the MMA N dimension (one pass over every input frame)
            if layer.p16_pack3d is None:
                from .tc_pack import pack_conv3d_fused_p16
                layer.p16_pack3d = pack_conv3d_fused_p16(layer.weight, layer.tc_scale16)
            call("pcab_conv3d_p16", P(s[0]), I(T), P(layer.p16_pack3d), F(1.0 / layer.tc_scale16), P(layer.bias), I(int(relu)), P(out),
                 I(n_img), I(H), I(W_), P(self._sat_counter(dev)), stream())
            path = "tc-p16"
        elif fmt:  # P16 activations in and out (csrc/conv_p16.cu)
            p0 = ctypes.c_void_p(s[0].data_ptr() + src0_off)
            call("pcab_conv3x3_p16", p0, I(c[0]), I(src0_cstride), P(s[1]), I(c[1]), P(s[2]), I(c[2]), Tl, P(layer.p16_pack),
                 F(1.0 / layer.tc_scale16), *tail, P(self._sat_counter(dev)), stream())
            path = "tc-p16"
        elif tc_ok and self.conv_operands == "f16":
            call("pcab_conv3x3_tc_f16", P(s[0]), I(c[0]), P(s[1]), I(c[1]), P(s[2]), I(c[2]), Tl, P(layer.tc_pack16),
                 F(1.0 / layer.tc_scale16), *tail, I(layer.cout), I(0), stream())
            path = "tc-f16pair"
        elif tc_ok:
            if layer.tc_pack is None:
                layer.tc_pack = self._pack_tc(layer)
            call("pcab_conv3x3_tc", P(s[0]), I(c[0]), P(s[1]), I(c[1]), P(s[2]), I(c[2]), Tl, P(layer.tc_pack), *tail,
                 I(layer.cout), I(0), stream())
            path = "tc"
        else:
            call("pcab_conv3x3_f32", P(s[0]), I(c[0]), P(s[1]), I(c[1]), P(s[2]), I(c[2]), Tl, P(layer.pack), *tail,
                 I(layer.cout), I(0), stream())
            path = "f32"
        if self.flop_acc is not None or ev is not None:
            if layer.temporal:
                taps = 9 * layer.splits[0] * (3 * T - 2) * (n_img // T)
            else:
                taps = 9 * sum(layer.splits) * n_img
            if self.flop_acc is not None:
                k = self._flop_key or "outside_graphs"
                acc = self.flop_acc.setdefault(k, {"flops": 0.0, "launches": 0, "bytes": 0.0})
                acc["flops"] += 2.0 * taps * layer.cout * H * W_
                acc["launches"] += 1
                cin_r = (sum(layer.splits) if not layer.temporal else layer.splits[0]) * n_img
                acc["bytes"] += 4.0 * H * W_ * (cin_r + layer.cout * n_img) + 2.0 * 9 * sum(layer.splits) * layer.cout * 2
        if ev is not None:
            e1.record()
            cin_read = (sum(layer.splits) if not layer.temporal else layer.splits[0]) * n_img
            nbytes = 4.0 * H * W_ * (cin_read + layer.cout * n_img) + 4.0 * 9 * sum(layer.splits) * layer.cout
            ev.append((e0, e1, 2.0 * taps * layer.cout * H * W_, path, nbytes))
        return out

    def _pack_tc(self, layer):
        from .tc_pack import pack_conv_tc
        return pack_conv_tc(layer)

    def _unet(self, W, prefix, x, n_img, H, W_, depth, final, fmt=0):
        enc = []
        h, w = H, W_
        dev = x.device
        for i in range(depth):
            x = self._conv(W[f"{prefix}d{i}c1"], [x], n_img, h, w, True, fmt=fmt)
            x = self._conv(W[f"{prefix}d{i}c2"], [x], n_img, h, w, True, fmt=fmt)
            enc.append((x, h, w))
            if i < depth - 1:
                c = x.shape[-1]
                pooled = torch.empty(n_img, h // 2, w // 2, c, device=dev, dtype=torch.float32)
                call("pcab_maxpool2x2", P(x), P(pooled), I(n_img), I(h), I(w), I(c), I(fmt), stream())
                x, h, w = pooled, h // 2, w // 2
        for i in range(depth - 1):
            skip, sh, sw = enc[-(i + 2)]
            up_l = W[f"{prefix}u{i}up"]
            pack, bias, cin, cout = up_l[:4]
            up = torch.empty(n_img, sh, sw, cout, device=dev, dtype=torch.float32)
            if fmt:  # ConvTranspose on the tensor cores (1-tap GEMM with 4 x Cout columns, scattered by the output maps)
                if up_l[4].get("p16") is None:
                    from .tc_pack import f16_weight_scale, pack_convT_p16
                    sc = f16_weight_scale(up_l[4]["weight"])
                    up_l[4]["p16"] = (pack_convT_p16(up_l[4]["weight"], sc), sc)
                wp, sc = up_l[4]["p16"]
                call("pcab_convT2x2_p16", P(x), I(cin), P(wp), F(1.0 / sc), P(bias), P(up), I(n_img), I(h), I(w), I(cout),
                     P(self._sat_counter(dev)), stream())
                if self.flop_acc is not None:
                    acc = self.flop_acc.setdefault(self._flop_key or "outside_graphs", {"flops": 0.0, "launches": 0, "bytes": 0.0})
                    acc["flops"] += 2.0 * 4 * cin * cout * n_img * h * w
                    acc["launches"] += 1
                    acc["bytes"] += 4.0 * n_img * h * w * (cin + 4 * cout) + 2.0 * 4 * cin * cout * 2
            else:
                call("pcab_convT2x2_f32", P(x), P(pack), P(bias), P(up), I(n_img), I(h), I(w), I(cin), I(cout), I(cout), I(0), stream())
            h, w = sh, sw
            x = self._conv(W[f"{prefix}u{i}c1"], [up, skip], n_img, h, w, True, fmt=fmt)
            x = self._conv(W[f"{prefix}u{i}c2"], [x], n_img, h, w, True, fmt=fmt)
        if final:
            x = self._conv(W[f"{prefix}final"], [x], n_img, h, w, False, fmt=fmt)
        return x

    # ------------------------------------------------------------------------------------------
    def _count(self, t):
        return int(t.item())

    def _mark(self, name):
        if self.stage_marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.stage_marks.append((name, e))

    def _select_async(self, n, dev, flags=None, values=None, value=0):
        """Like ``_select`` but the count travels to a pinned host buffer asynchronously: returns (idx buffer, wait) where
        ``wait()`` blocks until the count has arrived and returns (idx[:k], k).  Issued early, the wait is free."""
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        if n > 0:
            ws = scratch(size("pcab_select_workspace", I(n)), dev)
            call("pcab_select_indices", P(flags), P(values), L.L(value), I(n), P(idx), P(cnt), P(ws), Z(ws.numel()), stream())
        if getattr(self, "_pin_ints", None) is None:  # pinned staging for small readbacks (allocated once: cudaHostAlloc is slow)
            self._pin_ints, self._pin_next = torch.empty(16, dtype=torch.int32).pin_memory(), 0
        host = self._pin_ints[self._pin_next:self._pin_next + 1]
        self._pin_next = (self._pin_next + 1) % 16
        host.copy_(cnt, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def wait():
            ev.synchronize()
            k = int(host[0])
            return idx[:k], k

        return wait

    def _select(self, n, dev, flags=None, values=None, value=0):
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        if n > 0:
            ws = scratch(size("pcab_select_workspace", I(n)), dev)
            call("pcab_select_indices", P(flags), P(values), L.L(value), I(n), P(idx), P(cnt), P(ws), Z(ws.numel()), stream())
        k = self._count(cnt)
        return idx[:k], k

    # ------------------------------------------------------------------------------------------
    def forward(self, input_dict):
        """Same contract as ``models/motionnet.py:137-262``.  Inference only: the CUDA stages have no backward yet (SURVEY.md
        section 8 row f1 -- the loss and its gradients w.r.t. the outputs exist, ``pcaccumulation_b200/loss.py``), so a call
        that a training loop would differentiate (module in train() mode with autograd enabled) fails HERE instead of
        returning detached losses that silently never update the weights."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "pcaccumulation_b200.MotionNet has no backward pass: call it under torch.no_grad() / model.eval() "
                "(validation and test loops of libs/trainer.py and libs/tester.py); training needs the reference model")
        with torch.no_grad(), L.pinned_stream():
            return self._forward(input_dict)

    def _forward(self, input_dict):
        W = self._weights()
        st = self.stages = {}
        self._deferred = []  # (keys, device tensor): python floats of the API are fetched with ONE sync at the end
        cfg = self.cfg
        pts = input_dict["input_points"].float().contiguous()
        dev = pts.device
        if dev.type != "cuda":
            raise L.PcabError("pcaccumulation_b200.MotionNet runs on CUDA (sm_100a) only; move input_dict to the GPU")
        time_indice = input_dict.get("time_indice")  # absent on the SceneRunner fast path (int32 arrays in "_pcab")
        fb_labels = input_dict["fb_labels"]
        ego_gt = input_dict["ego_motion_gt"].float().contiguous()
        coordinates = input_dict.get("coordinates")
        num_voxels = input_dict["num_voxels"]
        shape = input_dict["shape"][0]
        Nx, Ny, nt = int(shape[0]), int(shape[1]), int(shape[3])
        M = input_dict["_pcab"]["coords_zyxt"].shape[0] if coordinates is None else coordinates.shape[0]
        N, B, T = pts.shape[0], num_voxels.shape[0], nt
        HW = Ny * Nx
        rng = host_floats(self.pc_range)
        vsz = host_floats(self.resolution)
        x_abs, y_abs = abs(float(self.pc_range[0])), abs(float(self.pc_range[1]))

        self._mark("start")
        # schema -> compact int32 device arrays (SceneRunner hands them over directly and skips the f64 round trip)
        fast = input_dict.get("_pcab")
        if fast is not None:
            p2v, pbatch, ptime = fast["p2v"], fast["pbatch"], fast["ptime"]
            coords_zyxt, pillar_batch = fast["coords_zyxt"], fast["pillar_batch"]
        else:
            p2v = input_dict["point_to_voxel_map"].reshape(-1).to(torch.int32).contiguous()
            pbatch = time_indice[:, 0].to(torch.int32).contiguous()
            ptime = time_indice[:, 1].to(torch.int32).contiguous()
            ci = coordinates.to(torch.int32)
            coords_zyxt = ci[:, 1:5].contiguous()
            pillar_batch = ci[:, 0].contiguous()
        pframe = (pbatch * T + ptime).contiguous() if B > 1 else ptime
        fb64 = fb_labels.reshape(-1).to(torch.int64).contiguous()

        # pillar index (stable sort by pillar) + statistics
        order = torch.empty(N, dtype=torch.int32, device=dev)
        pstart = torch.empty(M + 1, dtype=torch.int32, device=dev)
        ws = scratch(size("pcab_pillar_index_workspace", I(N)), dev)
        call("pcab_pillar_index", P(p2v), I(N), I(M), P(order), P(pstart), P(ws), Z(ws.numel()), stream())
        pillar_mean = torch.empty(M, 3, device=dev)
        fb_sub = torch.empty(M, dtype=torch.int32, device=dev)
        call("pcab_pillar_stats", P(pts), P(fb64), P(order), P(pstart), I(M), P(pillar_mean), P(fb_sub), stream())
        pillar_cell = torch.empty(M, dtype=torch.int32, device=dev)
        pillar_frame = torch.empty(M, dtype=torch.int32, device=dev)
        cell2pillar = torch.full((B * T * HW,), -1, dtype=torch.int32, device=dev)
        call("pcab_pillar_cells", P(coords_zyxt), P(pillar_batch), I(M), I(T), I(Ny), I(Nx), P(pillar_cell),
             P(pillar_frame), P(cell2pillar), stream())
        occ_map = torch.zeros(B, T, 1, Ny, Nx, device=dev)
        fb_map = torch.zeros(B, T, 1, Ny, Nx, dtype=torch.int64, device=dev)
        mean_map = torch.zeros(B * T, 3, Ny, Nx, device=dev)
        call("pcab_canvases", P(pillar_cell), P(fb_sub), P(pillar_mean), I(M), I(Ny), I(Nx), P(occ_map), P(fb_map),
             P(mean_map), stream())
        results = {"fb_seg_gt": fb_map, "occ_map": occ_map}

        self._mark("index+stats")
        # 1. pillar encoder -> BEV canvas
        fmt = self._fmt(B, T, Ny, Nx)  # activation format of the BEV tensors: 1 = P16 pairs, 0 = float32
        merged = bool(fmt and self.merge_heads)
        tc = (self.use_tensor_cores, self.conv_operands, fmt, self.merge_heads, self.fuse_conv3d)  # part of the graph key: a capture replays the kernels it recorded
        sat = self._sat_counter(dev)
        if fmt:
            sat.zero_()
        canvas = self._graph_input(("backbone", B, T, Ny, Nx, tc), (B * T, Ny, Nx, 32), dev)
        canvas.zero_()
        pillar_feats = torch.empty(M, 32, device=dev)
        ws = scratch(size("pcab_pillar_encode_workspace", I(N), I(M)), dev)
        if self.use_tensor_cores and self.pfn_tensor_cores:
            blob, bias_tc, inv9 = W["pfn_tc"]
            call("pcab_pillar_encode_tc", P(pts), P(ptime), P(order), P(p2v), P(coords_zyxt), P(pillar_cell), P(pillar_mean), P(blob),
                 P(bias_tc), inv9, I(N), I(M), rng, vsz, I(self.n_sweeps), P(pillar_feats), P(canvas), I(fmt), P(ws), Z(ws.numel()),
                 stream())
        else:
            call("pcab_pillar_encode", P(pts), P(ptime), P(order), P(p2v), P(pstart), P(coords_zyxt), P(pillar_cell),
                 P(pillar_mean), P(W["pfn"]), I(N), I(M), rng, vsz, I(self.n_sweeps), P(pillar_feats), P(canvas), I(fmt), P(ws),
                 Z(ws.numel()), stream())
        del ws

        self._mark("pillar_encoder")
        # 2. UNet backbone
        bev_feats, h = self._run_stack(("backbone", B, T, Ny, Nx, tc), self._backbone_stack(W, B, T, Ny, Nx, fmt))

        self._mark("unet")
        # 3. FG/BG head
        fb_seg = torch.empty(B, T, 2, Ny, Nx, device=dev)
        fb_est = torch.empty(B * T * HW, dtype=torch.int32, device=dev)
        call("pcab_head2_conv", P(h), I(32), I(h.shape[-1]), I(fmt), P(W["sem3"][0]), P(W["sem3"][1]), I(B * T), I(Ny), I(Nx),
             P(fb_seg), P(fb_est), stream())
        if "fb_est_map" in self.inject:  # [B,T,1,Ny,Nx] or flat: the argmax map the ego head and the gathers consume
            fb_est = self.inject["fb_est_map"].to(dev).reshape(-1).to(torch.int32).contiguous()
        fb_pp = torch.empty(N, 1, dtype=torch.int64, device=dev)
        call("pcab_fb_per_point", P(fb_est), P(pillar_cell), P(p2v), I(N), P(fb_pp), stream())
        results["fb_seg_est"] = fb_seg
        results["fb_est_per_points"] = fb_pp
        # foreground selection (models/motionnet.py:213-221): issued now, its count is read after the STPN stack is queued
        if self.mode in ("train", "val"):
            fg_flags = ((fb64 == 1) | (fb_pp[:, 0] == 1)).to(torch.int32)
            fg_wait = self._select_async(N, dev, flags=fg_flags)
        else:
            fg_wait = self._select_async(N, dev, values=fb_pp, value=1)

        self._mark("fb_head")
        # 4. ego-motion: the background-pillar counts start their trip to the host before the head convolutions are
        # queued, so the host draws the keypoint permutations while the GPU is busy with them
        prep = self._ego_prepare(cell2pillar, fb_est, M, B, T, Ny, Nx)
        if merged:  # channels 32..95 of the merged head tensor are the ego head's hidden layer
            geo = self._conv(W["ego3"], [h], B * T, Ny, Nx, False, fmt=fmt, src0_cstride=96, src0_off=128)
        else:
            h = self._conv(W["ego0"], [bev_feats], B * T, Ny, Nx, True, fmt=fmt)
            geo = self._conv(W["ego3"], [h], B * T, Ny, Nx, False, fmt=fmt)
        del h
        self._ego_motion(W, geo, cell2pillar, prep, pillar_mean, pillar_frame, M, ego_gt, B, T, Ny, Nx, results, fmt,
                         raw=(pts, pframe, fb_pp))
        if self.keep_stages:
            dec = (lambda t: t) if not fmt else _unpack_p16
            st.update(pillar_mean=pillar_mean, pillar_feats=pillar_feats, bev_feats=dec(bev_feats), geo=dec(geo), fb_est=fb_est)
        del geo

        self._mark("ego")
        # 5. warp + motion segmentation
        pose_est = results["ego_motion_est"].float().contiguous()
        if "ego_motion_est" in self.inject:
            pose_est = self.inject["ego_motion_est"].to(dev).float().contiguous()
        warped = self._graph_input(("stpn", B, T, Ny, Nx, tc), (B * T, Ny, Nx, 32), dev)
        call("pcab_warp_bev", P(bev_feats), P(pose_est), I(B), I(T), I(Ny), I(Nx), I(32), F(self.resolution[0]),
             F(self.resolution[1]), F(self.pc_range[0]), F(self.pc_range[1]), P(warped), I(fmt), stream())
        tp = torch.empty(N, 3, device=dev)
        call("pcab_transform_points", P(pts), P(pframe), P(pose_est), I(N), P(tp), stream())
        results["transformed_points"] = tp

        full_mos = torch.empty(N, 2, device=dev)
        full_off = torch.empty(N, 2, device=dev)
        call("pcab_init_point_outputs", I(N), P(full_mos), P(full_off), stream())
        # The Conv3d + STPN-UNet stack does not depend on WHICH points are foreground, only the guard `count > MIN_POINTS`
        # (models/motionnet.py:222) does: with a captured graph it is queued before the count is read (the readback then
        # costs nothing); its output is dropped in the rare scene that fails the guard.
        stack_key = ("stpn", B, T, Ny, Nx, tc)
        early = self.use_graphs and self.conv_events is None and self.flop_acc is None and \
            self._graphs.get(stack_key, {}).get("graph") is not None and self._graphs[stack_key].get("key") is self._packed_key
        mos_feats = self._run_stack(stack_key, self._stpn_stack(W, B, T, Ny, Nx, dev, fmt)) if early else None
        fg_idx, n_fg = fg_wait()
        if n_fg <= MIN_POINTS:
            mos_feats = None
        else:
            if not early:
                mos_feats = self._run_stack(stack_key, self._stpn_stack(W, B, T, Ny, Nx, dev, fmt))
            if self.use_tensor_cores:
                call("pcab_stpn_head_tc", P(mos_feats), I(fmt), I(Ny), I(Nx), P(tp), P(pbatch), P(fg_idx), I(n_fg), P(W["stpn_head_host"]),
                     P(W["stpn_head_tc1"]), P(W["stpn_head_tc"]), F(x_abs), F(y_abs), P(full_mos), P(full_off), stream())
            else:
                call("pcab_stpn_head", P(mos_feats), I(Ny), I(Nx), P(tp), P(pbatch), P(fg_idx), I(n_fg), P(W["stpn_head"]),
                     F(x_abs), F(y_abs), P(full_mos), P(full_off), stream())
        if self.keep_stages:
            dec = (lambda t: t) if not fmt else _unpack_p16
            st.update(warped=dec(warped), mos_feats=None if mos_feats is None else dec(mos_feats))
        results["mos_est"], results["offset_est"] = full_mos, full_off
        if "mos_est" in self.inject:
            full_mos = self.inject["mos_est"].to(dev).float().contiguous()
        if "offset_est" in self.inject:
            full_off = self.inject["offset_est"].to(dev).float().contiguous()
        if "transformed_points" in self.inject:
            tp = self.inject["transformed_points"].to(dev).float().contiguous()
        rec_est = tp.clone()
        results["rec_est"] = rec_est

        self._mark("warp+stpn")
        # 6. instances + TubeNet
        if self.mode in ("train", "val"):
            inst_labels = input_dict["inst_labels"][:, 0].long().contiguous()
            rec_idx, n_rec = self._select(N, dev, values=fb64, value=1)
        else:
            inst_labels, n_inst_max = self._cluster(tp, full_mos, full_off, input_dict["num_points"], B, N, dev)
            results["inst_labels_est"] = inst_labels
            self._mark("cluster")
            if "inst_labels_est" in self.inject:
                inst_labels = self.inject["inst_labels_est"].to(dev).long().contiguous()
                n_inst_max = int(inst_labels.max())
            rec_idx, n_rec = self._select(N, dev, flags=(inst_labels != 0).to(torch.int32))
        if n_rec > MIN_POINTS:
            if mos_feats is None:  # quirk Q4 (motionnet.py:222-245): upstream dies with NameError here
                raise NameError("name 'mos_feats' is not defined")
            bb = torch.empty(n_rec, 32, device=dev)
            mf = torch.empty(n_rec, 64, device=dev)
            call("pcab_ungrid", P(bev_feats), I(32), I(fmt), I(Ny), I(Nx), P(pts), P(pframe), P(rec_idx), I(n_rec), F(x_abs),
                 F(y_abs), P(bb), stream())
            call("pcab_ungrid", P(mos_feats), I(64), I(fmt), I(Ny), I(Nx), P(tp), P(pbatch), P(rec_idx), I(n_rec), F(x_abs),
                 F(y_abs), P(mf), stream())
            if self.keep_stages:
                st.update(backbone_feats=bb, motion_feats=mf)
            g_inst = torch.empty(n_rec, dtype=torch.int64, device=dev)
            g_batch = torch.empty(n_rec, dtype=torch.int64, device=dev)
            g_time = torch.empty(n_rec, dtype=torch.int64, device=dev)
            g_mos = torch.empty(n_rec, dtype=torch.int64, device=dev)
            g_tp = torch.empty(n_rec, 3, device=dev)
            sd64 = input_dict["sd_labels"].reshape(-1).to(torch.int64).contiguous()
            call("pcab_tpn_gather", P(rec_idx), I(n_rec), P(inst_labels), P(pbatch), P(ptime), P(tp), P(sd64), P(g_inst), P(g_batch),
                 P(g_time), P(g_tp), P(g_mos), stream())
            self._alignnet(W, {
                "inst_labels": g_inst, "batch_idx": g_batch, "time_idx": g_time,
                "n_inst_max": n_inst_max if self.mode == "test" else 0,
                "transformed_points": g_tp,
                "backbone_feats": bb, "motion_feats": mf, "inst_motion_gt": input_dict["inst_motion_gt"],
                "mos_labels": g_mos, "ego_motion_est": results["ego_motion_est"],
                "ego_motion_gt": results["ego_motion_gt"]}, results, T)
            call("pcab_scatter_rows3", P(results["sub_rec_est"]), P(rec_idx), I(n_rec), P(rec_est), stream())
        self._mark("tubenet")
        if fmt:
            self._deferred.append((("_p16_saturated",), sat))
        if self._deferred:
            vals = torch.cat([t.reshape(-1).float() for _, t in self._deferred]).cpu().tolist()
            i = 0
            for keys, _ in self._deferred:
                for k in keys:
                    results[k] = vals[i]
                    i += 1
        if results.pop("_p16_saturated", 0):
            # an activation left the fp16 range (+-65504) and was clamped: the fp16-pair operands cannot represent this
            # checkpoint's activations.  Switch this model to the 3xTF32 operands (full float32 range) and redo the scene.
            import warnings

            warnings.warn("pcaccumulation_b200: BEV activations beyond the fp16 range; switching conv_operands to 'tf32'")
            self.conv_operands = "tf32"
            return self._forward(input_dict)
        return results

    # ------------------------------------------------------------------------------------------
    def _ego_prepare(self, cell2pillar, fb_est, M, B, T, Ny, Nx):
        """Compact the occupied background cells per frame (cell order, like the boolean masks of
        models/egomotion.py:418-430) and start the asynchronous readback of the per-frame counts."""
        dev = cell2pillar.device
        nF = B * T
        ncell = nF * Ny * Nx
        bg_cells = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        frame_off = torch.empty(nF + 1, dtype=torch.int32, device=dev)
        ws = scratch(size("pcab_bg_compact_workspace", L.L(ncell)), dev)
        call("pcab_bg_compact", P(cell2pillar), P(fb_est), I(nF), I(Ny), I(Nx), P(bg_cells), P(frame_off), P(ws),
             Z(ws.numel()), stream())
        if getattr(self, "_pin_off", None) is None or self._pin_off.numel() != nF + 1:
            self._pin_off = torch.empty(nF + 1, dtype=torch.int32).pin_memory()
        host = self._pin_off
        host.copy_(frame_off, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return bg_cells, frame_off, host, ev

    def _ego_motion(self, W, geo, cell2pillar, prep, pillar_mean, pillar_frame, M, ego_gt, B, T, Ny, Nx, results, fmt=0, raw=None):
        """models/egomotion.py:387-469.  The host only draws the keypoint permutations (H3 protocol:
        ``torch.randperm`` on the CPU generator, in the reference's order) from ONE readback of the
        per-frame background-pillar counts; everything else is batched over all pairs on the device."""
        dev = geo.device
        cfg = self.cfg
        nF = B * T
        bg_cells, frame_off, host_off, ev = prep
        ev.synchronize()  # the one host wait of the ego head
        off = host_off.tolist()
        counts = [off[f + 1] - off[f] for f in range(nF)]
        mode = cfg["pose_estimation"]["seq_pose"]
        freq = cfg["data"]["freq"]
        pairs = []  # (batch, anchor, ref, duration)
        for b in range(B):
            if mode == "skip":
                pairs += [(b, 0, t, t / freq) for t in range(1, T)]
            elif mode == "chain":
                pairs += [(b, t - 1, t, 1.0 / freq) for t in range(1, T)]
            else:
                pairs += [(b, a, a + gap, gap / freq) for gap in range(1, T) for a in range(T - 1) if a + gap < T]
        npairs = len(pairs)

        def sample(n):
            if n <= 0:
                raise IndexError("no background pillars to register (models/egomotion.py:169)")
            if n > N_KPTS:
                return torch.randperm(n, generator=self.rng)[:N_KPTS]
            c = torch.arange(N_KPTS)
            c[n:] = n - 1
            return c

        # keypoint choices, pair table, squared distance gates and chain table travel in ONE pinned int32 buffer / ONE copy
        n_choice = npairs * 2 * N_KPTS
        n_words = n_choice + 2 * npairs + npairs + nF
        stage = getattr(self, "_ego_stage", None)
        if stage is None or stage[0].numel() != n_words or stage[1].device != dev:
            pin = torch.empty(n_words, dtype=torch.int32).pin_memory()
            stage = self._ego_stage = (pin, torch.empty(n_words, dtype=torch.int32, device=dev), pin.numpy())
        pin, dbuf, words = stage
        choice = words[:n_choice].reshape(npairs, 2, N_KPTS)
        pair_frames = words[n_choice:n_choice + 2 * npairs].reshape(npairs, 2)
        thr2 = words[n_choice + 2 * npairs:n_choice + 3 * npairs].view(np.float32)
        chain_pair = words[n_choice + 3 * npairs:]
        chain_pair[:] = -1
        for p, (b, anchor, ref, duration) in enumerate(pairs):
            choice[p, 0] = sample(counts[b * T + ref]).numpy()
            choice[p, 1] = sample(counts[b * T + anchor]).numpy()
            pair_frames[p, 0], pair_frames[p, 1] = b * T + ref, b * T + anchor
            thr2[p] = (duration * cfg["data"]["max_speed"]) ** 2
            if mode == "chain" or anchor == 0:
                chain_pair[b * T + ref] = p
        dbuf.copy_(pin, non_blocking=True)
        choice_d = dbuf[:n_choice]
        pair_frames_d = dbuf[n_choice:n_choice + 2 * npairs]
        thr2_d = dbuf[n_choice + 2 * npairs:n_choice + 3 * npairs]
        chain_pair_d = dbuf[n_choice + 3 * npairs:]
        perm = torch.empty(npairs, 1, N_KPTS, N_KPTS, device=dev)
        pose_pairs = torch.empty(npairs, 4, 4, device=dev)
        est = torch.empty(B, T, 4, 4, device=dev)
        gt = torch.empty(B, T, 4, 4, device=dev)
        scalars = torch.empty(4, device=dev)
        ws = scratch(size("pcab_ego_pairs_workspace", I(npairs)), dev)
        call("pcab_ego_pairs", P(geo), I(fmt), P(cell2pillar), P(pillar_mean), P(pillar_frame), I(M), P(bg_cells), P(frame_off),
             P(pair_frames_d), P(choice_d), P(thr2_d), I(npairs), P(W["alpha"]), P(W["beta"]),
             I(cfg["pose_estimation"]["sinkhorn_iter"]), P(ego_gt), P(chain_pair_d), I(B), I(T),
             I(1 if mode == "chain" else 0), P(perm), P(pose_pairs), P(est), P(gt), P(scalars), P(ws), Z(ws.numel()),
             stream())
        results["ego_l1_loss"], results["ego_l2_loss"] = scalars[0], scalars[1]
        if cfg["model"]["ego_icp"]:
            if self.keep_stages:
                self.stages["ego_motion_before_icp"] = est
            est = self._ego_icp(est, gt, raw, B, T, scalars)
        self._deferred.append((("ego_rot_error", "ego_trans_error"), scalars[2:4]))  # floats are read at the end of forward
        keep = [p for p, (b, anchor, ref, d) in enumerate(pairs) if mode == "chain" or anchor == 0]
        results["perm_matrix"] = [perm[p] for p in keep]
        results["ego_motion_est"], results["ego_motion_gt"] = est, gt
        if self.keep_stages:
            self.stages.update(pose_pairs=pose_pairs, choice=torch.from_numpy(choice.copy()), counts=counts)

    def _ego_icp(self, est, gt, raw, B, T, scalars):
        """models/egomotion.py:9-28,360-384,439-441 (model.ego_icp): every frame's raw background points are registered to
        the background points of frame 0 of their scene, starting from the estimated pose; the rotation / translation errors
        are those of the refined poses.  One batched call: problem = (scene, frame), target group = scene."""
        pts, pframe, fb_pp = raw
        dev = pts.device
        pe = self.cfg["pose_estimation"]
        bg = fb_pp[:, 0] == 0
        frame_t = pframe % T
        src_problem = torch.where(bg & (frame_t > 0), pframe, torch.full_like(pframe, -1)).to(torch.int32).contiguous()
        # targets: only the anchor frames' background points go into the search grid (a fifth of the cloud at T = 5)
        anchor = bg & (frame_t == 0)
        tgt = pts[anchor].contiguous()
        tgt_group = (pframe[anchor] // T).to(torch.int32).contiguous()
        problem_group = (torch.arange(B * T, device=dev, dtype=torch.int32) // T).contiguous()
        refined = torch.empty(B, T, 4, 4, device=dev)
        n, m = pts.shape[0], tgt.shape[0]
        ws = scratch(size("pcab_icp_workspace", I(m), I(B * T)), dev)
        call("pcab_icp_point_to_point", P(pts), P(src_problem), I(n), P(tgt), P(tgt_group), I(m), P(problem_group), I(B * T),
             P(est.contiguous()), F(pe["icp_threshold"]), I(pe["icp_max_iter"]), F(1e-6), F(1e-6), P(refined), P(None), P(ws),
             Z(ws.numel()), stream())
        call("pcab_ego_pose_errors", P(refined), P(gt), I(B), I(T), P(scalars[2:4]), stream())
        return refined

    def _tpn_icp(self, p_pts0, p_seg32, p_inst32, p_time32, final, K, T):
        """models/alignnet.py:54-112,264-266 (model.tpointnet_icp): per instance, the points of every later frame (moved by
        the regressed pose) are registered to the instance's frame-0 points; the result is composed onto the pose."""
        dev = p_pts0.device
        n = p_pts0.shape[0]
        rec = torch.empty_like(p_pts0)
        call("pcab_apply_seg_pose", P(p_pts0), P(p_seg32), P(final), I(n), P(rec), stream())
        src_problem = torch.where(p_time32 > 0, p_seg32, torch.full_like(p_seg32, -1)).contiguous()
        tgt_group = torch.where(p_time32 == 0, p_inst32, torch.full_like(p_inst32, -1)).contiguous()
        problem_group = (torch.arange(K * T, device=dev, dtype=torch.int32) // T).contiguous()
        upd = torch.empty(K, T, 4, 4, device=dev)
        ws = scratch(size("pcab_icp_workspace", I(n), I(K * T)), dev)
        call("pcab_icp_point_to_point", P(rec), P(src_problem), I(n), P(rec), P(tgt_group), I(n), P(problem_group), I(K * T),
             P(None), F(self.cfg["tpointnet"]["icp_threshold"]), I(50), F(1e-6), F(1e-6), P(upd), P(None), P(ws), Z(ws.numel()),
             stream())
        return torch.matmul(upd, final).contiguous()

    # ------------------------------------------------------------------------------------------
    def _cluster(self, tp, mos, off, num_points, B, N, dev):
        """models/cluster.py:86-110, per scene, fully on the device."""
        cc = self.cfg["cluster"]
        inst = torch.zeros(N, dtype=torch.int64, device=dev)
        npts = [int(v) for v in num_points.reshape(-1).tolist()]
        ninst_all = torch.zeros(max(B, 1), dtype=torch.int32, device=dev)
        n0 = 0
        for b in range(B):
            n = npts[b]
            if n <= 0:
                continue
            flags = torch.empty(n, dtype=torch.int32, device=dev)
            call("pcab_dynamic_flags", P(mos), I(n0), I(n), P(flags), stream())
            sel, s = self._select(n, dev, flags=flags)
            if s > cc["min_p_cluster"]:
                ws = scratch(size("pcab_cluster_workspace", I(s)), dev)
                call("pcab_cluster_scene", P(tp), P(off), P(sel), I(n0), I(s), F(0.05), D(cc["eps_dbscan"]),
                     I(cc["min_samples_dbscan"]), I(cc["min_p_cluster"]), P(inst), P(ninst_all[b:b + 1]), P(ws), Z(ws.numel()),
                     stream())
            n0 += n
        return inst, max(ninst_all.tolist())

    # ------------------------------------------------------------------------------------------
    def _alignnet(self, W, inp, results, T):
        """models/alignnet.py:166-285.  Relabelling / padding / row ordering run in two pcab calls
        (``pcab_tpn_relabel``, ``pcab_tpn_rows``) with one small readback between them; embeddings, regression and
        reconstruction are pcab kernels; only [K,T,4,4]-sized pose algebra stays in torch."""
        dev = inp["transformed_points"].device
        mos_labels = inp["mos_labels"]
        inst_labels = inp["inst_labels"]
        tb, t_idx = inp["batch_idx"], inp["time_idx"]
        tp = inp["transformed_points"].contiguous()
        n_points = inst_labels.size(0)
        ego_est, ego_gt = inp["ego_motion_est"], inp["ego_motion_gt"]
        test = self.mode == "test"
        if test:
            # upstream builds ONE identity motion list entry for the whole batch (alignnet.py:190-192), so every instance's
            # "GT" is the ego-pose error of scene 0 per frame and instance ids are not offset per scene
            K0 = inp["n_inst_max"] + 1
            G = torch.empty(T, 4, 4, device=dev)
            call("pcab_pose_error", P(ego_gt[0].contiguous()), P(ego_est[0].contiguous()), I(T), P(G), stream())
            motion_all = None
        else:
            inst_motion_gt = [m.to(dev).float() for m in inp["inst_motion_gt"]]
            upd = []
            for b, m in enumerate(inst_motion_gt):  # alignnet.py:9-38
                Kb = m.size(0)
                g = ego_gt[b][None].repeat(Kb, 1, 1, 1).view(-1, 4, 4)
                e = ego_est[b][None].repeat(Kb, 1, 1, 1).view(-1, 4, 4)
                upd.append((m.reshape(-1, 4, 4) @ g @ torch.linalg.inv(e)).view(Kb, -1, 4, 4))
            # alignnet.py:201-206: instance ids become global over the batch (scenes without points do not advance the offset)
            ks = torch.tensor([u.size(0) for u in upd], device=dev)
            nb = int(ego_est.shape[0])
            has = torch.bincount(tb, minlength=nb)[:len(upd)] > 0
            ks_eff = ks * has
            offs = torch.zeros(max(nb, len(upd)), dtype=torch.long, device=dev)
            offs[:len(upd)] = torch.cumsum(ks_eff, 0) - ks_eff
            inst_labels = inst_labels + offs[tb]
            motion_all = torch.cat(upd)
            K0 = motion_all.size(0)
        inst_labels = inst_labels.contiguous()
        t_idx = t_idx.contiguous()
        frame_count = torch.empty(K0 * T, dtype=torch.int32, device=dev)
        mapping = torch.empty(K0, dtype=torch.int32, device=dev)
        pad_frame = torch.empty(K0, dtype=torch.int32, device=dev)
        totals = torch.empty(2, dtype=torch.int32, device=dev)
        call("pcab_tpn_relabel", P(inst_labels), P(t_idx), I(n_points), I(K0), I(T), P(frame_count), P(mapping), P(pad_frame),
             P(totals), stream())
        K, n_extra = totals.tolist()  # the one readback of the TubeNet bookkeeping
        n_pad = n_points + n_extra
        seg32 = torch.empty(n_points, dtype=torch.int32, device=dev)
        inst_new = torch.empty(n_points, dtype=torch.int64, device=dev)
        p_idx32 = torch.empty(n_pad, dtype=torch.int32, device=dev)
        p_inst32 = torch.empty(n_pad, dtype=torch.int32, device=dev)
        p_time32 = torch.empty(n_pad, dtype=torch.int32, device=dev)
        p_seg32 = torch.empty(n_pad, dtype=torch.int32, device=dev)
        p_pts = torch.empty(n_pad, 3, device=dev)
        ws = scratch(size("pcab_tpn_rows_workspace", I(n_pad)), dev)
        call("pcab_tpn_rows", P(inst_labels), P(t_idx), I(n_points), I(n_extra), I(T), P(mapping), P(pad_frame), P(tp),
             P(seg32), P(inst_new), P(p_idx32), P(p_inst32), P(p_time32), P(p_seg32), P(p_pts), P(ws), Z(ws.numel()), stream())
        inst_labels = inst_new
        p_pts0 = p_pts  # padded_transformed_points_back (alignnet.py:226)
        if test:
            motion0_fn = lambda: G[None].expand(K, T, 4, 4).contiguous()
        else:
            motion_kept = motion_all[mapping >= 0]
            motion0_fn = lambda: motion_kept
        mos_emb = torch.empty(K, 128, device=dev)
        geo_emb = torch.empty(K, 128, device=dev)
        tc = self.use_tensor_cores
        if tc:
            for which, feat, (ws_tc, bias), dst in ((0, inp["motion_feats"], W["tpn_motion_tc"], mos_emb),
                                                    (1, inp["backbone_feats"], W["tpn_geo_tc"], geo_emb)):
                call("pcab_embed_segmax_tc", I(which), P(feat), P(p_idx32), P(p_inst32), I(n_pad), I(K), P(ws_tc[0]), P(ws_tc[1]),
                     P(ws_tc[2]), P(bias), P(dst), stream())
        else:
            call("pcab_tpn_static_embed", P(inp["motion_feats"]), P(inp["backbone_feats"]), P(p_idx32), P(p_inst32), I(n_pad),
                 I(K), P(W["tpn_motion"]), P(W["tpn_geo"]), P(mos_emb), P(geo_emb), stream())
        pos_tc, pos_bias = W["tpn_pos_tc"]
        pos_scratch = torch.empty(n_pad * 33, device=dev) if tc else None
        results["tpointnet_loss_terms"] = {}
        final = None
        ws = scratch(size("pcab_tpn_iteration_workspace", I(K), I(T)), dev)
        poses = []

        def gt_motion_at(it):
            """GT instance motion seen by iteration `it` (alignnet.py:250-254 applied for the earlier iterations)."""
            m = motion0_fn().reshape(-1, 4, 4).clone()
            for c in poses[:it]:
                c = c.reshape(-1, 4, 4)
                m[:, :3, :3] = torch.matmul(m[:, :3, :3], c[:, :3, :3].transpose(1, 2))
                m[:, :3, 3] = m[:, :3, 3] - torch.matmul(m[:, :3, :3], c[:, :3, 3].unsqueeze(-1)).squeeze(-1)
            return m.view(K, T, 4, 4)

        for it in range(self.cfg["tpointnet"]["n_iterations"]):
            pose = torch.empty(K, T, 4, 4, device=dev)
            pose_c = torch.empty(K * T, 4, 4, device=dev)
            rep = torch.empty(K * T, 7, device=dev)
            call("pcab_tpn_iteration", P(p_pts), P(p_inst32), P(p_time32), I(n_pad), I(K), I(T), P(mos_emb), P(geo_emb),
                 P(W["tpn_pos"]), P(W["tpn_reg"]), P(pose), P(pose_c), P(rep), P(ws), Z(ws.numel()),
                 P(pos_tc[0] if tc else None), P(pos_tc[1] if tc else None), P(pos_bias if tc else None), P(pos_scratch), stream())
            poses.append(pose)

            def compute(it=it, pts=p_pts, pose=pose, pose_c=pose_c, rep=rep):
                rows = p_idx32.long()
                return self._tpn_losses(pts, p_inst32.long(), p_time32.long(), p_seg32, mos_labels[rows], gt_motion_at(it),
                                        pose, pose_c, rep, K, T)

            terms = _LazyLossTerms(pose, compute)
            if not test:
                terms.materialize()  # the training / validation losses consume them
            results["tpointnet_loss_terms"][f"{it}_th"] = terms
            new_pts = torch.empty_like(p_pts)
            call("pcab_apply_seg_pose", P(p_pts), P(p_seg32), P(pose), I(n_pad), P(new_pts), stream())
            p_pts = new_pts
            c = pose.reshape(-1, 4, 4)
            final = c if final is None else torch.matmul(c, final)
        final = final.view(K, T, 4, 4).contiguous()
        if self.cfg["model"]["tpointnet_icp"]:
            final = self._tpn_icp(p_pts0, p_seg32, p_inst32, p_time32, final, K, T)
        rec_est = torch.empty(n_points, 3, device=dev)
        rec_gt = torch.empty(n_points, 3, device=dev)
        call("pcab_apply_seg_pose", P(tp), P(seg32), P(final), I(n_points), P(rec_est), stream())
        if test:
            t32 = t_idx.to(torch.int32)
            call("pcab_apply_seg_pose", P(tp), P(t32), P(G), I(n_points), P(rec_gt), stream())
        else:
            call("pcab_apply_seg_pose", P(tp), P(seg32), P(motion_kept.contiguous()), I(n_points), P(rec_gt), stream())
        errs = torch.empty(2, device=dev)
        acc4 = torch.empty(4, dtype=torch.float64, device=dev)
        call("pcab_inst_errors", P(rec_est), P(rec_gt), P(t_idx), P(mos_labels.contiguous()), I(n_points), P(acc4), P(errs), stream())
        self._deferred.append((("inst_l2_error", "dynamic_inst_l2_error"), errs))
        results["inst_labels_adjusted"] = inst_labels
        results["inst_pose_est"] = final
        results["sub_rec_est"] = rec_est

    def _tpn_losses(self, pts, inst, tidx, seg32, mos_labels, motion_gt, pose, pose_c, rep, K, T):
        """Loss terms of models/tpointnet.py:224-237,275-288 (bookkeeping on K*T rows; names swapped upstream, Q6)."""
        dev = pts.device
        seg = seg32.long()
        ones = torch.ones(seg.numel(), device=dev)
        frame_count = torch.zeros(K * T, device=dev).scatter_add_(0, seg, ones)
        fw = (frame_count > self.cfg["tpointnet"]["min_points"]).float()
        inst_mos = torch.zeros(K * T, dtype=mos_labels.dtype, device=dev).scatter_reduce_(0, seg, mos_labels, "amax", include_self=False)
        # quirk: upstream builds the weights with ones_like(<int64 labels>) and assigns 0.2 into that INTEGER tensor
        # (models/tpointnet.py:232-233), which truncates to 0 -- frames without a dynamic point get weight 0, not 0.2
        mw = (inst_mos != 0).float()
        tw = ((torch.arange(self.n_sweeps, device=dev) + 1).repeat(K) / self.n_sweeps).float()
        fw = fw * mw * tw
        sums = torch.zeros(K * T, 3, device=dev, dtype=torch.float64).index_add_(0, seg, pts.double())
        cen = (sums / frame_count.clamp(min=1).double()[:, None]).float()[::T]  # anchor-frame centroid per instance
        cen_r = cen.repeat_interleave(T, 0).unsqueeze(2)
        gt = motion_gt.reshape(-1, 4, 4).clone()
        gt[:, :3, 3] += torch.matmul(gt[:, :3, :3] - torch.eye(3, device=dev)[None], cen_r).squeeze(2)
        gt_quat = _mat2quat_scipy(gt[:, :3, :3])
        centered = pts - cen[inst]
        rec_e = torch.empty_like(centered)
        rec_g = torch.empty_like(centered)
        call("pcab_apply_seg_pose", P(centered.contiguous()), P(seg32), P(pose_c), I(seg.numel()), P(rec_e), stream())
        call("pcab_apply_seg_pose", P(centered.contiguous()), P(seg32), P(gt.contiguous()), I(seg.numel()), P(rec_g), stream())
        diff = rec_e - rec_g
        cnt = frame_count.clamp(min=1)
        f_l1 = torch.zeros(K * T, device=dev).scatter_add_(0, seg, torch.norm(diff, p=2, dim=1)) / cnt
        f_l2 = torch.zeros(K * T, device=dev).scatter_add_(0, seg, torch.norm(diff, p=1, dim=1)) / cnt
        wsum = fw.sum() + 1e-20
        quat = torch.nn.functional.normalize(rep[:, :4], p=2, dim=1)
        return {
            "l1_loss": (f_l1 * fw).sum() / wsum,
            "l2_loss": (f_l2 * fw).sum() / wsum,
            "rot_loss": (torch.norm(gt_quat - quat, p=2, dim=1) * fw).sum() / wsum,
            "trans_loss": (torch.norm(gt[:, :3, 3] - rep[:, 4:], p=2, dim=1) * fw).sum() / wsum,
            "inst_est_motion": pose,
        }


def _unpack_p16(t):
    from .tc_pack import unpack_p16

    return unpack_p16(t)


class _LazyLossTerms(dict):
    """Per-iteration TubeNet outputs (models/tpointnet.py:297-303).  ``inst_est_motion`` is always present; the four
    loss scalars (``l1_loss, l2_loss, rot_loss, trans_loss``) are only consumed by the training loss, so in test mode
    they are computed on first access instead of on the hot path."""

    _LOSS_KEYS = ("l1_loss", "l2_loss", "rot_loss", "trans_loss")

    def __init__(self, pose, compute):
        super().__init__(inst_est_motion=pose)
        self._compute = compute

    def materialize(self):
        if self._compute is not None:
            fn, self._compute = self._compute, None
            self.update(fn())
        return self

    def __missing__(self, key):
        if key in self._LOSS_KEYS and self._compute is not None:
            return self.materialize()[key]
        raise KeyError(key)

    def keys(self):
        self.materialize()
        return dict.keys(self)

    def items(self):
        self.materialize()
        return dict.items(self)

    def __contains__(self, key):
        return key in self._LOSS_KEYS or dict.__contains__(self, key)


def _mat2quat_scipy(R):
    """scipy.spatial.transform.Rotation.from_matrix(R).as_quat() (xyzw), as used at models/tpointnet.py:66-67:
    the input is projected onto SO(3) with an SVD (scipy does so for inputs that are not orthogonal to 1e-12, which
    is every float32 matrix; for exactly orthogonal ones the projection is the identity to 1e-16), then scipy's
    largest-component branch is taken.  Branch-free (no boolean-mask indexing) so it costs a handful of launches."""
    M = R.double()
    U, _, Vt = torch.linalg.svd(M)
    M = U @ Vt
    m = lambda i, j: M[:, i, j]
    d0, d1, d2 = m(0, 0), m(1, 1), m(2, 2)
    tr = d0 + d1 + d2
    dec = torch.stack((d0, d1, d2, tr), 1)
    choice = dec.argmax(1)
    q0 = torch.stack((1 - tr + 2 * d0, m(1, 0) + m(0, 1), m(2, 0) + m(0, 2), m(2, 1) - m(1, 2)), 1)
    q1 = torch.stack((m(0, 1) + m(1, 0), 1 - tr + 2 * d1, m(2, 1) + m(1, 2), m(0, 2) - m(2, 0)), 1)
    q2 = torch.stack((m(0, 2) + m(2, 0), m(1, 2) + m(2, 1), 1 - tr + 2 * d2, m(1, 0) - m(0, 1)), 1)
    q3 = torch.stack((m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1), 1 + tr), 1)
    allq = torch.stack((q0, q1, q2, q3), 1)  # [n, case, 4]
    q = allq.gather(1, choice.view(-1, 1, 1).expand(-1, 1, 4))[:, 0]
    q = q / torch.norm(q, dim=1, keepdim=True)
    return q.float()
