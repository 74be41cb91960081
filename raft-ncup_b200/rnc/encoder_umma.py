"""fnet / cnet (core/extractor.py:118-192 BasicEncoder, ResidualBlock :6-56) on the tensor-core convolution path.

The wide 3x3 / 1x1 layers run on rnc_conv2d_umma_fwd (fp16 hi/lo split operands, stride 1/2); so does the 7x7/2 stem: the
normalised image is repacked once as a zero-padded [H][W+8][4] plane of split halves and the convolution reads it through a
sliding-window tensor map (16-pixel windows 16 bytes apart: the TMA unit builds the im2col rows; 7 row taps x 64 = K 448 with
zero weights for the 9 phantom pixels and the phantom channel); InstanceNorm (fnet) is a statistics pass + an apply pass fused with
ReLU / residual add / re-splitting; BatchNorm (cnet, eval mode) is folded into the convolution weights.  The encoders write
their results straight into the loop's resident buffers: fmap1 -> f1_cl, fmap2 -> level 0 of f2_pyr, tanh(net) -> h and
hx[:, 0:128], relu(inp) -> hx[:, 128:256] (raft_nc_dbl.py:129-140).
"""
import ctypes as C

import torch

from . import native
from .engine import _ptr, _stream
from .engine_umma import HX_LD, SplitBuf, UmmaWeights

EPS = 1e-5


def _fold_bn(conv, bn):
    w, b = conv.weight.detach().float(), conv.bias.detach().float()
    if bn is None:
        return w, b
    s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    return w * s.view(-1, 1, 1, 1), (b - bn.running_mean) * s + bn.bias.detach()


class PackedEncoder:
    """Kernel-ready weights of one BasicEncoder.  kind = 'instance' (fnet) or 'batch' (cnet, BN folded)."""

    def __init__(self, enc):
        self.kind = enc.norm_fn
        if self.kind not in ("instance", "batch"):
            raise NotImplementedError("tensor-core encoder supports the reference's two configurations: instance / batch norm")
        bn = self.kind == "batch"
        if bn and any(m.training for m in enc.modules() if isinstance(m, torch.nn.BatchNorm2d)):
            raise NotImplementedError("cnet BatchNorm in training mode (batch statistics) is not built; call .eval() / freeze_bn()")
        w, b = _fold_bn(enc.conv1, enc.norm1 if bn else None)
        # window form of the 7x7x3 filter: input "channel" e = 4 * px + c of the 16-pixel window, one tap per filter row
        wv = torch.zeros(w.shape[0], 16, 4, 7, dtype=torch.float32, device=w.device)
        wv[:, :7, :3, :] = w.permute(0, 3, 1, 2)              # [o][kx][c][ky]
        self.stem = UmmaWeights(wv.reshape(w.shape[0], 64, 7, 1), b, [64])
        self.blocks = []
        for layer in (enc.layer1, enc.layer2, enc.layer3):
            for blk in layer:
                cin, cout = blk.conv1.in_channels, blk.conv1.out_channels
                stride = blk.conv1.stride[0]
                w1 = UmmaWeights(*_fold_bn(blk.conv1, blk.norm1 if bn else None), [cin])
                w2 = UmmaWeights(*_fold_bn(blk.conv2, blk.norm2 if bn else None), [cout])
                wd = None
                if blk.downsample is not None:
                    wd = UmmaWeights(*_fold_bn(blk.downsample[0], blk.norm3 if bn else None), [cin])
                self.blocks.append((cin, cout, stride, w1, w2, wd))
        self.head = UmmaWeights(enc.conv2.weight, enc.conv2.bias, [128])


class EncoderBuffers:
    """Scratch for one encoder pass over N images of Hin x Win (three resolution levels)."""

    def __init__(self, device, N, Hin, Win):
        self.key = (str(device), N, Hin, Win)
        f = dict(dtype=torch.float32, device=device)
        self.dims = []
        h, w = (Hin + 1) // 2, (Win + 1) // 2
        for c in (64, 96, 128):
            self.dims.append((h, w, c))
            h, w = (h + 1) // 2, (w + 1) // 2
        self.X32, self.XS, self.T32, self.AS, self.D32 = [], [], [], [], []
        for (h, w, c) in self.dims:
            rows = N * h * w
            self.X32.append(torch.empty(rows, c, **f))
            self.T32.append(torch.empty(rows, c, **f))
            self.D32.append(torch.empty(rows, c, **f) if c != 64 else None)
            self.XS.append(SplitBuf(rows, c, device))
            self.AS.append(SplitBuf(rows, c, device))
        # stem input: zero-padded pixel plane [N][Hin][pitch][4] of split halves (+ 32 zero pixels: the last windows run over)
        self.pitch = (Win + 7) & ~1
        npx = N * Hin * self.pitch + 32
        self.img_hi = torch.zeros(npx, 4, dtype=torch.float16, device=device)
        self.img_lo = torch.zeros(npx, 4, dtype=torch.float16, device=device)
        self.stats = torch.zeros(N * 128 * 2, dtype=torch.float64, device=device)   # kept zeroed by rnc_instnorm_finalize
        self.mr = torch.empty(N * 128 * 2, **f)


class EncoderRunner:
    def __init__(self, engine):
        self.eng = engine
        self.L = engine.L
        self._bufs = None

    def packed(self, enc):
        return self.eng._packed_for("enc", enc, PackedEncoder)

    def buffers(self, device, N, Hin, Win):
        if self._bufs is None or self._bufs.key != (str(device), N, Hin, Win):
            self._bufs = None
            self._bufs = EncoderBuffers(device, N, Hin, Win)
        return self._bufs

    # ------------------------------------------------------------------ instance-norm helpers
    def _norm(self, bufs, x32, N, P, Cc, mode, res=None, out32=None, split=None, fused_stats=False):
        """fused_stats: the producing convolution already accumulated the sums into bufs.stats (rnc_conv_umma_desc.stats)."""
        s = _stream()
        if fused_stats:
            native.check(self.L.rnc_instnorm_finalize(_ptr(bufs.stats), N, P, Cc, EPS, _ptr(bufs.mr), s), "instnorm_finalize")
        else:
            native.check(self.L.rnc_instnorm_stats(_ptr(x32), N, P, Cc, EPS, _ptr(bufs.stats), _ptr(bufs.mr), s), "instnorm_stats")
        native.check(self.L.rnc_instnorm_apply(_ptr(x32), _ptr(bufs.mr), _ptr(res), N, P, Cc, mode, _ptr(out32),
                                               C.c_void_p(split.hi.data_ptr() if split else 0),
                                               C.c_void_p(split.lo.data_ptr() if split else 0), s), "instnorm_apply")

    def _trunk(self, pk, bufs, image, N, Hin, Win):
        """Stem + the six residual blocks.  Leaves the 128-channel features at 1/8 resolution in bufs.XS[2] (split)."""
        E, eng, s = native, self.eng, _stream()
        inst = pk.kind == "instance"
        h, w, _ = bufs.dims[0]
        native.check(self.L.rnc_stem_window_prep(_ptr(image), N, Hin, Win, bufs.pitch, _ptr(bufs.img_hi), _ptr(bufs.img_lo), s),
                     "stem_window_prep")
        win = dict(stride=2, hin=Hin, win=w, win_pitch=4 * bufs.pitch, flags=eng.conv_flags | E.CONV_WINDOW)
        img = (bufs.img_hi.data_ptr(), bufs.img_lo.data_ptr())
        if inst:
            eng.uconv(N, h, w, img, 64, 8, pk.stem, E.EPI_LINEAR, out_f32=bufs.T32[0].data_ptr(), ldo_f32=64,
                      stats=bufs.stats.data_ptr(), **win)
            self._norm(bufs, bufs.T32[0], N, h * w, 64, 1, out32=bufs.X32[0], split=bufs.XS[0], fused_stats=True)
        else:
            eng.uconv(N, h, w, img, 64, 8, pk.stem, E.EPI_RELU, out_f32=bufs.X32[0].data_ptr(), ldo_f32=64,
                      out_split=bufs.XS[0].ptrs(), ldo_split=64, **win)
        lvl = 0
        for bi, (cin, cout, stride, w1, w2, wd) in enumerate(pk.blocks):
            # the fp32 copy of a block's output is only read as the next block's residual (blocks without a downsample branch)
            need32 = bi + 1 < len(pk.blocks) and pk.blocks[bi + 1][5] is None
            src = lvl
            if stride == 2:
                lvl += 1
            hi_, wi_, _ = bufs.dims[src]
            h, w, _ = bufs.dims[lvl]
            P = h * w
            xs_in, x32_in = bufs.XS[src], bufs.X32[src]
            if inst:
                st = bufs.stats.data_ptr()
                eng.uconv(N, h, w, xs_in.ptrs(), cin, cin, w1, E.EPI_LINEAR, out_f32=bufs.T32[lvl].data_ptr(), ldo_f32=cout,
                          stride=stride, hin=hi_, win=wi_, stats=st)
                self._norm(bufs, bufs.T32[lvl], N, P, cout, 1, split=bufs.AS[lvl], fused_stats=True)
                res = x32_in
                if wd is not None:
                    eng.uconv(N, h, w, xs_in.ptrs(), cin, cin, wd, E.EPI_LINEAR, out_f32=bufs.T32[lvl].data_ptr(), ldo_f32=cout,
                              stride=stride, hin=hi_, win=wi_, stats=st)
                    self._norm(bufs, bufs.T32[lvl], N, P, cout, 0, out32=bufs.D32[lvl], fused_stats=True)
                    res = bufs.D32[lvl]
                eng.uconv(N, h, w, bufs.AS[lvl].ptrs(), cout, cout, w2, E.EPI_LINEAR, out_f32=bufs.T32[lvl].data_ptr(), ldo_f32=cout,
                          stats=st)
                self._norm(bufs, bufs.T32[lvl], N, P, cout, 2, res=res, out32=bufs.X32[lvl] if need32 else None, split=bufs.XS[lvl],
                           fused_stats=True)
            else:
                eng.uconv(N, h, w, xs_in.ptrs(), cin, cin, w1, E.EPI_RELU, out_split=bufs.AS[lvl].ptrs(), ldo_split=cout,
                          stride=stride, hin=hi_, win=wi_)
                res = x32_in
                if wd is not None:
                    eng.uconv(N, h, w, xs_in.ptrs(), cin, cin, wd, E.EPI_LINEAR, out_f32=bufs.D32[lvl].data_ptr(), ldo_f32=cout,
                              stride=stride, hin=hi_, win=wi_)
                    res = bufs.D32[lvl]
                eng.uconv(N, h, w, bufs.AS[lvl].ptrs(), cout, cout, w2, E.EPI_RELU_ADD_RELU,
                          out_f32=bufs.X32[lvl].data_ptr() if need32 else 0, ldo_f32=cout, out_split=bufs.XS[lvl].ptrs(),
                          ldo_split=cout, res=res.data_ptr(), ldres=cout)
        return bufs.dims[2]

    # ------------------------------------------------------------------ public
    def run(self, model, ws, image1, image2):
        """image1/image2: raw [B,3,H,W] fp32 in 0..255 (the stem normalises).  Fills ws.f1_cl, ws.f2_pyr (level 0), ws.h,
        ws.hx[:, :256]; the caller finishes the pyramid."""
        eng, E = self.eng, native
        B, _, Hin, Win = image1.shape
        dev = image1.device
        pf, pc = self.packed(model.fnet), self.packed(model.cnet)
        bufs = self.buffers(dev, 2 * B, Hin, Win)
        # ---- fnet on both frames (extractor.py:168-172 concatenates them along the batch)
        both = torch.cat([image1, image2], 0).contiguous()
        h8, w8, _ = self._trunk(pf, bufs, both, 2 * B, Hin, Win)
        P = h8 * w8
        eng.alloc_fmaps(ws, B, 256, h8, w8, 4)
        xs = bufs.XS[2]
        eng.uconv(B, h8, w8, xs.ptrs(), 128, 128, pf.head, E.EPI_LINEAR, out_f32=ws.f1_cl.data_ptr(), ldo_f32=256)
        off = B * P * 128 * 2                                  # second half of the batch inside the split planes (bytes)
        eng.uconv(B, h8, w8, (xs.hi.data_ptr() + off, xs.lo.data_ptr() + off), 128, 128, pf.head, E.EPI_LINEAR,
                  out_f32=ws.f2_pyr.data_ptr(), ldo_f32=256)
        # ---- cnet on frame 1
        self._trunk(pc, bufs, image1.contiguous(), B, Hin, Win)
        ws.gru_const_valid = False                              # inp changes: the GRU's hoisted share must be recomputed
        eng.uconv(B, h8, w8, bufs.XS[2].ptrs(), 128, 128, pc.head, E.EPI_TANH_RELU, out_f32=ws.h.data_ptr(), ldo_f32=128,
                  out_split=ws.hx.ptrs(), ldo_split=HX_LD)
        return h8, w8
