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
from .engine import CORR_CH, _ptr, _require_cuda, _stream, make_engine

# --------------------------------------------------------------------------------------------- encoders (C6)
# Run once per pair, outside the per-iteration path: kept on cuDNN in strict fp32 (SURVEY.md §8f-1 "next").


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
        _require_cuda(fmap1, fmap2)
        self.num_levels, self.radius = num_levels, radius
        self.engine = engine or make_engine()
        B, D, H, W = fmap1.shape
        self.ws = _LookupState(B, D, H, W)
        self.engine.fmap_prepare(self.ws, fmap1.detach().float().contiguous(), fmap2.detach().float().contiguous(), num_levels)

    def __call__(self, coords):
        _require_cuda(coords)
        ws = self.ws
        side = 2 * self.radius + 1
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
    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)


class SepConvGRU(nn.Module):
    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(hidden_dim + input_dim, hidden_dim, k, padding=p))


class BasicMotionEncoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 256, 1, padding=0)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)


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
        self._engine = None

    def engine(self):
        if self._engine is None:
            self._engine = make_engine()
        return self._engine

    def forward(self, net, inp, corr, flow, upsample=True):
        _require_cuda(net, inp, corr, flow)
        if torch.is_grad_enabled() and any(t.requires_grad for t in (net, inp, corr, flow)):
            raise NotImplementedError("backward through the fused update block is not built yet (SURVEY.md §8f-3)")
        eng = self.engine()
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
    """Parameter holder for core/nconv_modules.py:140-215: stores ``weight_p``; the effective kernel is
    softplus(weight_p, beta=10) (EnforcePos, :218-269)."""

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


class NConvUNet(nn.Module):
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


class Simple(nn.Module):
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
        self._engine = None

    def engine(self):
        if self._engine is None:
            self._engine = make_engine()
        return self._engine

    def forward(self, x_lowres, x_guidance=None, out_scale=1.0):
        _require_cuda(x_lowres, x_guidance)
        if torch.is_grad_enabled() and (x_lowres.requires_grad or x_guidance.requires_grad):
            raise NotImplementedError("backward through the fused NCUP kernel is not built yet (SURVEY.md §8f-3)")
        if any(isinstance(m, nn.BatchNorm2d) and m.training for m in self.weights_est_net.modules()):
            raise NotImplementedError("weights-net BatchNorm in training mode (batch statistics) is not built; call .eval() / freeze_bn()")
        B, C, h, w = x_lowres.shape
        if C != 2 or x_guidance.shape[1] != 128 or h != 2 * x_guidance.shape[2] or w != 2 * x_guidance.shape[3]:
            raise ValueError("expected x_lowres [B,2,h,w] with guidance [B,128,h/2,w/2]")
        eng = self.engine()
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
