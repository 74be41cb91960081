"""The reference's module surface (same class names, ctor args, forward signatures, state_dict keys), with every
hot-path forward routed to librnc.so through rnc.engine.  Parameter containers are ordinary nn.Modules created in
the reference's construction order, so ``torch.manual_seed(s); RAFT(args)`` yields the reference's exact weights.

Reference surfaces mirrored here (under /root/reference/core):
  corr.py:6-55 CorrBlock        update.py:6-141 FlowHead/SepConvGRU/BasicMotionEncoder/BasicUpdateBlock
  extractor.py:6-56,118-192 ResidualBlock/BasicEncoder      interp_weights_est.py:10-47 Simple
  nconv_modules.py:25-215 NConvUNet/NConv2d                 upsampler.py:10-210 get_upsampler/NConvUpsampler
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native
from .engine import (CORR_CH, HX_LD, PackedFlowHead, PackedGRU, PackedMotionEncoder, _ptr, _require_cuda, _stream, engine_for,
                     module_tensors, pack_conv, pack_thin)

# --------------------------------------------------------------------------------------------- encoders (C6)
# RAFT.forward runs the encoders on the tensor-core path (rnc/encoder_umma.py).  The nn.Module forwards below are the
# reference's own layer graph on cuDNN in strict fp32: used only when a caller invokes fnet/cnet directly, with
# RNC_ENCODER=cudnn, or under args.mixed_precision.


def _grad_needed(module, *tensors):
    return torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in tensors)
                                        or any(p.requires_grad for p in module_tensors(module)))


class _Seam:
    """Common prologue of the operator seams: one CUDA device for all inputs, that device's engine, its lock."""

    def __init__(self, *tensors):
        self.dev = _require_cuda(*tensors)
        self.eng = engine_for(self.dev)
        self._guard = torch.cuda.device(self.dev)

    def __enter__(self):
        self._guard.__enter__()
        self.eng.lock.acquire()
        return self.eng

    def __exit__(self, *exc):
        self.eng.lock.release()
        return self._guard.__exit__(*exc)


def _make_norm(kind, ch):
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "group":
        return nn.GroupNorm(num_groups=ch // 8, num_channels=ch)
    if kind == "none":
        return nn.Sequential()
    raise ValueError(kind)


class ResidualBlock(nn.Module):
    def __init__(self, in_planes, planes, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = _make_norm(norm_fn, planes), _make_norm(norm_fn, planes)
        self.downsample = None
        if stride != 1:
            self.norm3 = _make_norm(norm_fn, planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        skip = x if self.downsample is None else self.downsample(x)
        return self.relu(skip + y)


class BasicEncoder(nn.Module):
    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = nn.GroupNorm(8, 64) if norm_fn == "group" else _make_norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        widths = [(64, 64, 1), (64, 96, 2), (96, 128, 2)]
        for i, (cin, cout, stride) in enumerate(widths, 1):
            setattr(self, f"layer{i}", nn.Sequential(ResidualBlock(cin, cout, norm_fn, stride),
                                                     ResidualBlock(cout, cout, norm_fn, 1)))
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        pair = isinstance(x, (tuple, list))
        if pair:
            n = x[0].shape[0]
            x = torch.cat(x, 0)
        # strict fp32: TF32 convolutions alone move the final flow by ~1e-2 px
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            x = self.relu1(self.norm1(self.conv1(x)))
            x = self.layer3(self.layer2(self.layer1(x)))
            x = self.conv2(x)
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, [n, n], 0) if pair else x


# --------------------------------------------------------------------------------------------- CorrBlock (A1-A3)


class CorrBlock:
    """Drop-in for core/corr.py:6-44.  The constructor stores CL feature maps and the pooled fmap2 pyramid (never the
    4-D volume); ``__call__(coords)`` returns the reference's [N, levels*(2r+1)^2, H, W] fp32 tensor."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, engine=None):
        self.dev = _require_cuda(fmap1, fmap2)
        self.num_levels, self.radius = num_levels, radius
        self.engine = engine or engine_for(self.dev)
        B, D, H, W = fmap1.shape
        self.ws = _LookupState(B, D, H, W)
        self.fmap1, self.fmap2 = fmap1, fmap2            # kept for the training path (gradients flow to the feature maps)
        with torch.cuda.device(self.dev), self.engine.lock:
            self.engine.fmap_prepare(self.ws, fmap1.detach().float().contiguous(), fmap2.detach().float().contiguous(), num_levels)

    def __call__(self, coords):
        if _require_cuda(coords) != self.dev:
            raise ValueError("coords must live on the feature maps' device")
        if torch.is_grad_enabled() and (self.fmap1.requires_grad or self.fmap2.requires_grad):
            from .train import corr_lookup_autograd
            return corr_lookup_autograd(self, coords)
        ws = self.ws
        side = 2 * self.radius + 1
        with torch.cuda.device(self.dev), self.engine.lock:
            out = torch.empty(ws.B, self.num_levels * side * side, ws.H8, ws.W8, dtype=torch.float32, device=coords.device)
            self.engine.lookup(ws, coords.detach().float().contiguous(), out, 0, 0, self.radius)
        return out

    @staticmethod
    def corr(fmap1, fmap2):
        raise NotImplementedError("the all-pairs volume (corr.py:47-55) is never materialised by this implementation")


class _LookupState:
    def __init__(self, B, D, H, W):
        self.B, self.D, self.H8, self.W8 = B, D, H, W
        self.f1_cl = self.f2_pyr = None
        self.levels = 4


# --------------------------------------------------------------------------------------------- update block (A5-A8)


class FlowHead(nn.Module):
    """core/update.py:6-14: conv2(relu(conv1(x))), x NCHW [B,128,H,W] -> [B,2,H,W]."""

    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)

    def forward(self, x):
        if _grad_needed(self, x):
            from .train import flow_head_train
            return flow_head_train(self, x)
        if x.shape[1] != 128 or self.conv1.out_channels != 256:
            raise NotImplementedError("kernels are built for the reference's FlowHead(128, 256)")
        with _Seam(x) as eng:
            B, _, H, W = x.shape
            ws = eng.ffma_workspace(x.device, B, H, W)
            pk = eng._packed_for("flow_head", self, PackedFlowHead)
            native.check(eng.L.rnc_nchw_to_cl(_ptr(x.detach().float().contiguous()), B, 128, H, W, _ptr(ws.hx), HX_LD, 0, _stream()),
                         "nchw_to_cl")
            eng._flow_head_ffma(ws, pk, want_delta=True)       # also advances the workspace's scratch coords1 (unused here)
            return ws.delta.clone()


class SepConvGRU(nn.Module):
    """core/update.py:33-60: forward(h [B,128,H,W], x [B,256,H,W]) -> h."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(hidden_dim + input_dim, hidden_dim, k, padding=p))

    def forward(self, h, x):
        if _grad_needed(self, h, x):
            from .train import sep_conv_gru_train
            return sep_conv_gru_train(self, h, x)
        if h.shape[1] != 128 or x.shape[1] != 256:
            raise NotImplementedError("kernels are built for the reference's SepConvGRU(128, 256)")
        with _Seam(h, x) as eng:
            B, _, H, W = h.shape
            ws = eng.ffma_workspace(h.device, B, H, W)
            pk = eng._packed_for("gru", self, PackedGRU)
            s = _stream()
            native.check(eng.L.rnc_nchw_to_cl(_ptr(h.detach().float().contiguous()), B, 128, H, W, _ptr(ws.hx), HX_LD, 0, s), "nchw_to_cl")
            native.check(eng.L.rnc_nchw_to_cl(_ptr(x.detach().float().contiguous()), B, 256, H, W, _ptr(ws.hx), HX_LD, 128, s), "nchw_to_cl")
            eng._gru_ffma(ws, pk)
            return Engine_net_nchw(eng, ws)


def Engine_net_nchw(eng, ws):
    out = torch.empty(ws.B, 128, ws.H8, ws.W8, dtype=torch.float32, device=ws.hx.device)
    native.check(eng.L.rnc_cl_to_nchw(_ptr(ws.hx), HX_LD, 0, ws.B, 128, ws.H8, ws.W8, _ptr(out), _stream()), "cl_to_nchw")
    return out


class BasicMotionEncoder(nn.Module):
    """core/update.py:79-97: forward(flow [B,2,H,W], corr [B,324,H,W]) -> cat([conv features (126), flow (2)])."""

    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 256, 1, padding=0)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        if _grad_needed(self, flow, corr):
            from .train import motion_encoder_train
            return motion_encoder_train(self, flow, corr)
        if corr.shape[1] != CORR_CH:
            raise NotImplementedError("kernels are built for 4 levels x radius 4 = 324 correlation channels")
        with _Seam(flow, corr) as eng:
            B, _, H, W = flow.shape
            ws = eng.ffma_workspace(flow.device, B, H, W)
            pk = eng._packed_for("motion_encoder", self, PackedMotionEncoder)
            s = _stream()
            native.check(eng.L.rnc_nchw_to_cl(_ptr(corr.detach().float().contiguous()), B, CORR_CH, H, W, _ptr(ws.corr), CORR_CH, 0, s),
                         "nchw_to_cl(corr)")
            native.check(eng.L.rnc_coords_init(_ptr(ws.coords1), _ptr(flow.detach().float().contiguous()), B, H, W, s), "coords_init")
            eng._motion_encoder_ffma(ws, pk)
            out = torch.empty(B, 128, H, W, dtype=torch.float32, device=flow.device)
            native.check(eng.L.rnc_cl_to_nchw(_ptr(ws.hx), HX_LD, 256, B, 128, H, W, _ptr(out), s), "cl_to_nchw")
            return out


class BasicUpdateBlock(nn.Module):
    """Drop-in for core/update.py:114-141: forward(net, inp, corr, flow) -> (net, mask, delta_flow), all NCHW."""

    def __init__(self, args, hidden_dim=128, input_dim=128):
        super().__init__()
        self.args = args
        if hidden_dim != 128 or args.corr_levels != 4 or args.corr_radius != 4:
            raise NotImplementedError("kernels are built for the reference's only live config: hidden 128, 4 levels, radius 4")
        self.encoder = BasicMotionEncoder(args)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1, padding=0))
        self.net = []

    def engine(self, device=None):
        from .engine import module_device
        return engine_for(device if device is not None else module_device(self))

    def forward(self, net, inp, corr, flow, upsample=True):
        if _grad_needed(self, net, inp, corr, flow):
            from .train import update_block_train
            return update_block_train(self, net, inp, corr, flow)
        with _Seam(net, inp, corr, flow) as eng:
            return self._forward(eng, net, inp, corr, flow)

    def _forward(self, eng, net, inp, corr, flow):
        B, _, H, W = net.shape
        pk = eng.packed_update(self)
        ws = eng.workspace(net.device, B, H, W, pk.has_mask, False)
        s = _stream()
        L = eng.L
        eng.load_state(ws, net.float().contiguous(), inp.float().contiguous())
        eng.load_corr(ws, corr.float().contiguous())
        # the kernels read flow as coords1 - grid: rebuild coords1 from the flow argument
        native.check(L.rnc_coords_init(_ptr(ws.coords1), _ptr(flow.float().contiguous()), B, H, W, s), "coords_init")
        eng.update_iter(ws, pk, want_mask=pk.has_mask, want_delta=True)
        net_out = eng.net_nchw(ws)
        self.net = net_out                                  # guidance tap read by raft_nc_dbl.py:161
        mask = None
        if pk.has_mask:
            mask = torch.empty(B, 576, H, W, dtype=torch.float32, device=net.device)
            native.check(L.rnc_cl_to_nchw(_ptr(ws.mask), 576, 0, B, 576, H, W, _ptr(mask), s), "cl_to_nchw(mask)")
        else:
            mask = 0.25 * net_out                           # `.25 * Sequential()(net)` of the reference (update.py:140)
        return net_out, mask, ws.delta.clone()


# --------------------------------------------------------------------------------------------- NCUP (U2-U7)


class NConv2d(nn.Module):
    """core/nconv_modules.py:140-215: stores ``weight_p``; the effective kernel is softplus(weight_p, beta=10) (EnforcePos,
    :218-269), recomputed at every forward.  forward((data, conf)) -> (nconv, conf_out), NCHW, through rnc_nconv2d_fwd."""

    def __init__(self, in_channels, out_channels, kernel_size, pos_fn="softplus", bias=False):
        super().__init__()
        if bias or pos_fn.lower() != "softplus":
            raise NotImplementedError("only the shipped NConv config (SoftPlus, no bias) is built")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, tuple(kernel_size)
        self.eps = 1e-20
        w = torch.empty(out_channels, in_channels, *self.kernel_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))         # _ConvNd.reset_parameters: consumed, then overwritten (:207-209)
        n = self.kernel_size[0] * self.kernel_size[1] * out_channels
        w.normal_(2, math.sqrt(2.0 / n))
        self.weight_p = nn.Parameter(F.softplus(w, beta=10))

    @property
    def weight(self):
        return F.softplus(self.weight_p, beta=10)

    def forward(self, inpt):
        data, conf = inpt[0], inpt[1]
        if _grad_needed(self, data, conf):
            from .train import nconv2d_autograd
            return nconv2d_autograd(data, conf, self.weight, self.eps)
        return nconv2d_forward(data, conf, self.weight.detach(), self.eps)


def nconv2d_forward(data, conf, weight, eps=1e-20):
    """One normalized convolution through the C ABI (nconv_modules.py:164-199); weight = the positive kernel."""
    with _Seam(data, conf, weight) as eng:
        N, Cin, H, W = data.shape
        Cout, _, kh, kw = weight.shape
        if conf.shape != data.shape or weight.shape[1] != Cin:
            raise ValueError("NConv2d: data/conf/weight shapes do not match")
        y = torch.empty(N, Cout, H, W, dtype=torch.float32, device=data.device)
        c = torch.empty_like(y)
        native.check(eng.L.rnc_nconv2d_fwd(_ptr(data.detach().float().contiguous()), _ptr(conf.detach().float().contiguous()),
                                           _ptr(weight.detach().float().contiguous()), N, Cin, Cout, H, W, kh, kw, eps,
                                           _ptr(y), _ptr(c), _stream()), "nconv2d")
        return y, c


class NConvUNet(nn.Module):
    """core/nconv_modules.py:25-136 at the configuration every reference script ships.  forward((data, conf)) ->
    (xout, cout): at num_downsampling = 1 the decoder consumes x[1] twice (index quirk at :128-131), so the pooled branch
    never reaches the output and only nconv_in -> nconv_x2[0] -> decoder[0](cat(x1, x1)) -> nconv_out is computed."""

    def __init__(self, in_ch=1, channels_multiplier=2, num_downsampling=1, encoder_filter_sz=5, decoder_filter_sz=3,
                 out_filter_sz=1, pos_fn="SoftPlus", groups=1, use_bias=False, data_pooling="conf_based",
                 shared_encoder=True, use_double_conv=False):
        super().__init__()
        self.__name__ = "NConvUNet"
        if (in_ch, channels_multiplier, num_downsampling, encoder_filter_sz, decoder_filter_sz, out_filter_sz,
                shared_encoder, use_double_conv, use_bias, groups) != (1, 2, 1, 5, 3, 1, True, False, False, 1):
            raise NotImplementedError("NCUP kernel is built for the configuration every reference script ships (SURVEY.md §5)")
        c = in_ch * channels_multiplier
        self.num_downsampling, self.data_pooling = num_downsampling, data_pooling
        self.nconv_in = NConv2d(in_ch, c, (5, 5), pos_fn)
        self.nconv_x2 = nn.Sequential(NConv2d(c, c, (5, 5), pos_fn))
        self.encoder = nn.ModuleList([nn.Sequential(self.nconv_in, self.nconv_x2), self.nconv_x2[0]])
        self.decoder = nn.ModuleList([NConv2d(2 * c, c, (3, 3), pos_fn)])
        self.nconv_out = NConv2d(c, in_ch, (1, 1), pos_fn)

    def forward(self, inpt):
        x, c = self.nconv_in((inpt[0], inpt[1]))
        x, c = self.nconv_x2[0]((x, c))
        x, c = self.decoder[0]((torch.cat((x, x), 1), torch.cat((c, c), 1)))
        return self.nconv_out((x, c))


class Simple(nn.Module):
    """core/interp_weights_est.py:10-47: forward(x [B,130,h,w]) -> final_act(out(conv[1](conv[0](x))))."""

    def __init__(self, num_ch, out_ch, filter_sz, dilation=None, final_act=torch.sigmoid, use_bn=False):
        super().__init__()
        self.__name__ = "Simple"
        if list(filter_sz) != [3, 3, 1] or (dilation is not None and any(d != 1 for d in dilation)) or len(num_ch) != 3:
            raise NotImplementedError("weights net kernels are built for filter_sz [3,3,1], dilation 1 (SURVEY.md §5)")
        self.in_ch, self.num_layers = num_ch[0], len(num_ch) - 1
        self.conv = nn.ModuleList()
        for i in range(self.num_layers):
            layers = [nn.Conv2d(num_ch[i], num_ch[i + 1], 3, padding=1)]
            if use_bn:
                layers.append(nn.BatchNorm2d(num_ch[i + 1]))
            layers.append(nn.ReLU(inplace=True))
            self.conv.append(nn.Sequential(*layers))
        self.out = nn.Conv2d(num_ch[-1], out_ch, 1)
        self.final_act = final_act

    def forward(self, x):
        if _grad_needed(self, x):
            from .train import simple_train
            return simple_train(self, x)
        if self.final_act is not torch.sigmoid or self.out.out_channels != 2:
            raise NotImplementedError("the fused confidence head applies the sigmoid the reference wires in (upsampler.py:44-46)")
        if any(isinstance(m, nn.BatchNorm2d) and m.training for m in self.modules()):
            raise NotImplementedError("weights-net BatchNorm with batch statistics needs the training path (enable grad)")
        with _Seam(x) as eng:
            from .engine import pack_conv as _pc
            B, Cin, h, w = x.shape
            cpad = (Cin + 3) // 4 * 4
            folded = eng._packed_for("simple", self, lambda m: _PackedSimple(m, cpad))
            s = _stream()
            M = B * h * w
            xin = torch.zeros(M, cpad, dtype=torch.float32, device=x.device) if cpad != Cin else torch.empty(M, cpad, dtype=torch.float32, device=x.device)
            native.check(eng.L.rnc_nchw_to_cl(_ptr(x.detach().float().contiguous()), B, Cin, h, w, _ptr(xin), cpad, 0, s), "nchw_to_cl")
            c0, c1 = folded.c_mid
            g1 = torch.empty(M, (c0 + 3) // 4 * 4, dtype=torch.float32, device=x.device)
            g2 = torch.empty(M, (c1 + 3) // 4 * 4, dtype=torch.float32, device=x.device)
            eng.conv(B, h, w, xin.data_ptr(), cpad, cpad, folded.g0, c0, 3, 3, native.EPI_RELU, g1.data_ptr(), g1.shape[1])
            eng.conv(B, h, w, g1.data_ptr(), c0, g1.shape[1], folded.g1, c1, 3, 3, native.EPI_RELU, g2.data_ptr(), g2.shape[1])
            conf = torch.empty(B, 2, h, w, dtype=torch.float32, device=x.device)
            native.check(eng.L.rnc_conf_head_fwd(_ptr(g2), c1, g2.shape[1], _ptr(folded.gout[0]), _ptr(folded.gout[1]), B, h, w,
                                                 _ptr(conf), s), "conf_head")
            return conf


def fold_simple_convs(wn):
    """(weight, bias) of Simple's two 3x3 layers with eval-mode BatchNorm folded in (interp_weights_est.py:26-30)."""
    convs = []
    for blk in wn.conv:
        conv = blk[0]
        w, b = conv.weight.detach().float(), conv.bias.detach().float()
        if len(blk) == 3:   # Conv, BatchNorm, ReLU
            bn = blk[1]
            sc = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            w = w * sc.view(-1, 1, 1, 1)
            b = (b - bn.running_mean) * sc + bn.bias.detach()
        convs.append((w, b))
    return convs


class _PackedSimple:
    def __init__(self, wn, cin_pad):
        convs = fold_simple_convs(wn)
        self.g0 = pack_conv(convs[0][0], convs[0][1], cin_pad=cin_pad)
        self.g1 = pack_conv(convs[1][0], convs[1][1], cin_pad=(convs[1][0].shape[1] + 3) // 4 * 4)
        self.c_mid = (convs[0][0].shape[0], convs[1][0].shape[0])
        self.gout = (pack_thin(wn.out.weight), wn.out.bias.detach().float().contiguous())


class NConvUpsampler(nn.Module):
    """Drop-in for core/upsampler.py:75-210.  forward(x_lowres [B,2,h,w], x_guidance [B,128,h/2,w/2]) -> [B,2,4h,4w]."""

    def __init__(self, scale=None, size=None, interpolation_net=None, weights_est_net=None, use_data_for_guidance=True,
                 channels_to_batch=True, use_residuals=False, est_on_high_res=False):
        super().__init__()
        self.__name__ = "NConvUpsampler"
        if scale is None and size is None:
            raise ValueError("Either scale or size needs to be set!")
        if scale is not None and size is not None:
            raise ValueError("You can set either scale or size at a time!")
        if interpolation_net is None:
            raise ValueError("An interpolation network mush be provided!")
        if scale != 4 or not use_data_for_guidance or not channels_to_batch or use_residuals or est_on_high_res \
                or weights_est_net is None:
            raise NotImplementedError("NCUP kernel is built for scale 4 / data-for-guidance / channels-to-batch (SURVEY.md §5)")
        self.scaleH = self.scaleW = float(scale)
        self.interpolation_net, self.weights_est_net = interpolation_net, weights_est_net
        self.use_data_for_guidance, self.channels_to_batch = use_data_for_guidance, channels_to_batch
        self.use_residuals, self.est_on_high_res = use_residuals, est_on_high_res

    def engine(self, device=None):
        from .engine import module_device
        return engine_for(device if device is not None else module_device(self))

    def forward(self, x_lowres, x_guidance=None, out_scale=1.0):
        if _grad_needed(self, x_lowres, x_guidance):
            from .train import ncup_upsampler_train
            return ncup_upsampler_train(self, x_lowres, x_guidance, out_scale)
        if any(isinstance(m, nn.BatchNorm2d) and m.training for m in self.weights_est_net.modules()):
            raise NotImplementedError("weights-net BatchNorm with batch statistics needs the training path (enable grad)")
        B, C, h, w = x_lowres.shape
        if C != 2 or x_guidance is None or x_guidance.shape[1] != 128 or h != 2 * x_guidance.shape[2] or w != 2 * x_guidance.shape[3]:
            raise ValueError("expected x_lowres [B,2,h,w] with guidance [B,128,h/2,w/2]")
        with _Seam(x_lowres, x_guidance) as eng:
            pu = eng.packed_upsampler(self)
            ws = eng.workspace(x_lowres.device, B, h // 2, w // 2, False, True)
            g_cl = torch.empty(B * (h // 2) * (w // 2), 128, dtype=torch.float32, device=x_lowres.device)
            native.check(eng.L.rnc_nchw_to_cl(_ptr(x_guidance.float().contiguous()), B, 128, h // 2, w // 2, _ptr(g_cl), 128, 0,
                                              _stream()), "nchw_to_cl(guidance)")
            return eng.ncup_from_lowres(ws, pu, x_lowres.float().contiguous(), g_cl.data_ptr(), 128, out_scale)


def get_upsampler(in_ch, guidance_ch, args):
    """core/upsampler.py:10-72 — the factory is hard-wired to the NConv upsampler (:12)."""
    interpolation_net = NConvUNet(in_ch=1, channels_multiplier=args.interp_net_channels_multiplier,
                                  num_downsampling=args.interp_net_num_downsampling,
                                  encoder_filter_sz=args.interp_net_encoder_filter_sz,
                                  decoder_filter_sz=args.interp_net_decoder_filter_sz,
                                  out_filter_sz=args.interp_net_out_filter_sz, use_bias=args.interp_net_use_bias,
                                  data_pooling=args.interp_net_data_pooling, shared_encoder=args.interp_net_shared_encoder,
                                  use_double_conv=args.interp_net_use_double_conv, pos_fn="SoftPlus", groups=1)
    num_channels = list(args.weights_est_net_num_ch)
    num_channels.insert(0, guidance_ch + in_ch if args.final_upsampling_use_data_for_guidance else guidance_ch)
    use_bn = args.dataset == "sintel"                        # upsampler.py:42
    if args.weights_est_net.lower() != "simple":
        raise NotImplementedError("only the `Simple` weights-estimation net is built (every reference script selects it)")
    weights_est_net = Simple(num_ch=num_channels, out_ch=in_ch, use_bn=use_bn, filter_sz=args.weights_est_net_filter_sz,
                             dilation=args.weights_est_net_dilation, final_act=torch.sigmoid)
    return NConvUpsampler(scale=args.final_upsampling_scale, interpolation_net=interpolation_net,
                          weights_est_net=weights_est_net, use_data_for_guidance=args.final_upsampling_use_data_for_guidance,
                          channels_to_batch=args.final_upsampling_channels_to_batch,
                          use_residuals=args.final_upsampling_use_residuals, est_on_high_res=args.final_upsampling_est_on_high_res)
