"""RAFT / RAFT-NCUP model shells: the drop-in boundary.

Mirrors ``RAFT(nn.Module)`` of core/raft.py:24-143 (convex upsampler) and core/raft_nc_dbl.py:26-173 (NCUP upsampler):
same ctor Namespace, ``forward(image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False)``, return values
and state_dict keys.  The iteration loop runs entirely on resident channel-last buffers through librnc.so.
"""
import ctypes as C
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import native
from .engine import _Timed, _ptr, _require_cuda, _stream, engine_for, module_device, module_tensors
from .modules import BasicEncoder, BasicUpdateBlock, get_upsampler


class _RAFTBase(nn.Module):
    ncup = False

    def __init__(self, args):
        super().__init__()
        self.args = args
        if args.small:
            # the reference's --small path crashes (raft.py:134 passes align_corners= to upflow8)
            raise NotImplementedError("--small is out of scope: it is broken in the reference and unused by its scripts")
        self.hidden_dim = self.context_dim = 128
        args.corr_levels = 4
        args.corr_radius = 4
        args.dropout = 0     # raft.py:41-42: `'dropout' not in args._get_kwargs()` is always true -> dropout forced to 0
        self.fnet = BasicEncoder(output_dim=256, norm_fn="instance", dropout=args.dropout)
        self.cnet = BasicEncoder(output_dim=256, norm_fn="batch", dropout=args.dropout)
        self.update_block = BasicUpdateBlock(self.args, hidden_dim=128)

    # ------------------------------------------------------------------ reference helpers kept verbatim in meaning
    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def initialize_flow(self, img):
        """raft.py:63-71: two identical coordinate grids at 1/8 resolution (flow = coords1 - coords0)."""
        N, _, H, W = img.shape
        ys, xs = torch.meshgrid(torch.arange(H // 8, device=img.device), torch.arange(W // 8, device=img.device), indexing="ij")
        g = torch.stack([xs, ys], 0).float()[None].repeat(N, 1, 1, 1)
        return g, g.clone()

    def engine(self, device=None):
        """The per-device engine (rnc.engine.engine_for) this model's forwards run on.  Nothing is cached on the module:
        it stays deep-copyable / picklable, and nn.DataParallel replicas resolve their own device's engine."""
        return engine_for(device if device is not None else module_device(self))

    # ------------------------------------------------------------------ forward
    def _needs_grad(self):
        return torch.is_grad_enabled() and any(t.requires_grad for t in module_tensors(self))

    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        """Estimate optical flow between a pair of frames (raft_nc_dbl.py:115-173 / raft.py:87-143)."""
        dev = _require_cuda(image1, image2, flow_init)
        pdev = module_device(self)
        if pdev != dev:
            raise ValueError(f"model parameters are on {pdev} but the images are on {dev}")
        eng = engine_for(dev)
        # the kernels launch on `dev`'s current stream whatever device is current in the calling thread (nn.DataParallel
        # worker threads, a model on cuda:1 in a cuda:0 process); one forward at a time per device
        with torch.cuda.device(dev), eng.lock:
            return self._forward(eng, image1, image2, iters, flow_init, test_mode)

    def _forward(self, eng, image1, image2, iters, flow_init, test_mode):
        if iters < 1:
            raise ValueError("iters must be >= 1")
        if hasattr(self, "data_idx"):
            self.data_idx += 1
        if self._needs_grad():
            # training path (train.py:215): the same graph with autograd, exact-fp32 kernels forward and backward
            if image1.shape[2] % 8 or image1.shape[3] % 8:
                raise ValueError("image height/width must be multiples of 8 (pad with utils.utils.InputPadder, evaluate.py:125)")
            from .train import raft_forward_train
            return raft_forward_train(self, image1, image2, iters, flow_init, test_mode)
        B, _, Him, Wim = image1.shape
        if Him % 8 or Wim % 8:
            raise ValueError("image height/width must be multiples of 8 (pad with utils.utils.InputPadder, evaluate.py:125)")
        if test_mode and eng.graphs_enabled(self):
            return eng.graph_forward(self, image1, image2, iters, flow_init)
        return self._forward_eager(eng, image1, image2, iters, flow_init, test_mode)

    def _forward_eager(self, eng, image1, image2, iters, flow_init, test_mode):
        B, _, Him, Wim = image1.shape
        H8, W8 = Him // 8, Wim // 8
        L = eng.L
        pk = eng.packed_update(self.update_block)
        pu = eng.packed_upsampler(self.upsampler) if self.ncup else None
        if self.ncup and any(isinstance(m, nn.BatchNorm2d) and m.training for m in self.upsampler.modules()):
            raise NotImplementedError("weights-net BatchNorm with batch statistics needs the training path (enable grad) "
                                      "or eval mode: call .eval() / freeze_bn()")
        ws = eng.workspace(image1.device, B, H8, W8, pk.has_mask, self.ncup)
        s = _stream()
        amp = bool(getattr(self.args, "mixed_precision", False))
        if eng.mode == "umma" and not amp and os.environ.get("RNC_ENCODER", "umma").lower() == "umma":
            # encoders on the tensor-core path, writing straight into the resident buffers (raft_nc_dbl.py:118-140)
            with _Timed(eng, "encoders"):
                eng.encoder().run(self, ws, image1.float().contiguous(), image2.float().contiguous())
                eng.finish_fmaps(ws)
        else:
            image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
            image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
            with torch.autocast("cuda", enabled=amp):
                fmap1, fmap2 = self.fnet([image1, image2])
            fmap1, fmap2 = fmap1.float().contiguous(), fmap2.float().contiguous()
            with torch.autocast("cuda", enabled=amp):
                cnet = self.cnet(image1)
                net, inp = torch.split(cnet, [128, 128], dim=1)
                net, inp = torch.tanh(net), torch.relu(inp)
            net, inp = net.float().contiguous(), inp.float().contiguous()
            eng.fmap_prepare(ws, fmap1, fmap2, 4)
            eng.load_state(ws, net, inp)
        fi = None
        if flow_init is not None:
            fi = flow_init.to(image1.device).float().contiguous()
            if fi.shape != (B, 2, H8, W8):
                raise ValueError("flow_init must be [N,2,H/8,W/8]")
        native.check(L.rnc_coords_init(_ptr(ws.coords1), _ptr(fi), B, H8, W8, s), "coords_init")

        preds = []
        flow_up = None
        for itr in range(iters):
            last = itr == iters - 1
            need_up = last or not test_mode     # inference upsamples once (SURVEY.md finding 9); list mode needs all
            eng.begin_iter(ws, pk)      # forks convf1 (CUDA cores) onto a side stream; it co-runs with the tensor-core lookup
            eng.lookup_resident(ws)
            eng.update_iter(ws, pk, want_mask=(pk.has_mask and need_up))
            if need_up:
                flow_up = self._upsample(eng, ws, pu)
                preds.append(flow_up)
        self.update_block.net = eng.net_nchw(ws)
        if test_mode:
            return eng.flow_low(ws), flow_up
        return preds

    def _upsample(self, eng, ws, pu):
        raise NotImplementedError


class RAFTConvex(_RAFTBase):
    """core/raft.py:24-143 — baseline RAFT with the convex-combination upsampler."""
    ncup = False

    def _upsample(self, eng, ws, pu):
        return eng.convex_upsample(ws, eng.flow_low(ws), ws.mask, 576)

    def upsample_flow(self, flow, mask):
        """raft.py:73-84: flow [N,2,H8,W8], mask [N,576,H8,W8] (NCHW) -> [N,2,8*H8,8*W8]."""
        dev = _require_cuda(flow, mask)
        eng = engine_for(dev)
        B, _, H8, W8 = flow.shape
        with torch.cuda.device(dev), eng.lock:
            m_cl = torch.empty(B * H8 * W8, 576, dtype=torch.float32, device=dev)
            native.check(eng.L.rnc_nchw_to_cl(_ptr(mask.detach().float().contiguous()), B, 576, H8, W8, _ptr(m_cl), 576, 0, _stream()),
                         "nchw_to_cl(mask)")
            ws = _Dims(B, H8, W8)
            return eng.convex_upsample(ws, flow.detach().float().contiguous(), m_cl, 576)


class RAFTNcup(_RAFTBase):
    """core/raft_nc_dbl.py:26-173 — RAFT with the NCUP (normalized-convolution) upsampler."""
    ncup = True

    def __init__(self, args):
        super().__init__(args)
        if getattr(args, "load_pretrained", None) is not None:
            # raft_nc_dbl.py:58-66: strict load of a baseline-RAFT checkpoint (keys carry the DataParallel `module.` prefix)
            state = torch.load(args.load_pretrained, map_location="cpu")
            self.load_state_dict(OrderedDict((k[7:], v) for k, v in state.items()))
        self.update_block.mask = nn.Sequential()             # raft_nc_dbl.py:68
        if getattr(args, "freeze_raft", False):
            for p in self.parameters():
                p.requires_grad = False
        self.upsampler = get_upsampler(2, 128, args)
        self.data_idx = 0

    def _upsample(self, eng, ws, pu):
        native.check(eng.L.rnc_flow_x2_fwd(_ptr(ws.coords1), ws.B, ws.H8, ws.W8, _ptr(ws.x4), _stream()), "flow_x2")
        gptr, gld = eng.guidance(ws)
        return eng.ncup_from_lowres(ws, pu, ws.x4, gptr, gld, 8.0)   # `8 *` of raft_nc_dbl.py:161

    def upsample_flow(self, flow_lr, guidance):
        """raft_nc_dbl.py:107-112 (without the caller's x8): nearest x2, then the NConv upsampler."""
        x4 = torch.nn.functional.interpolate(flow_lr, scale_factor=2, mode="nearest")
        return self.upsampler(x4, guidance)


class _Dims:
    def __init__(self, B, H8, W8):
        self.B, self.H8, self.W8 = B, H8, W8
