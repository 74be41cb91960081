"""256-column layers: one 256-wide item per pixel tile vs two 128-wide items (flags bit RNC_CONV_SPLIT_N)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from rnc import native
from rnc.engine_umma import UmmaEngine, SplitBuf, UmmaWeights
eng = UmmaEngine(); dev = "cuda:0"
B, H, W = int(os.environ.get("B", 8)), 55, 128
for (cin, cout, kh, kw) in ((352, 256, 1, 1), (384, 256, 1, 5), (384, 256, 5, 1), (128, 256, 3, 3), (256, 256, 3, 3)):
    g = torch.Generator().manual_seed(0)
    buf = SplitBuf(B * H * W, cin, dev); buf.hi.normal_(); buf.lo.normal_(0, 1e-3)
    wt = UmmaWeights(torch.randn(cout, cin, kh, kw, generator=g).to(dev) / 40, torch.zeros(cout, device=dev), [cin])
    out = SplitBuf(B * H * W, wt.coutpad, dev)
    for fl in (0, 4):
        for _ in range(3):
            eng.uconv(B, H, W, buf.ptrs(), cin, cin, wt, native.EPI_RELU, out_split=out.ptrs(), ldo_split=wt.coutpad, flags=fl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.uconv(B, H, W, buf.ptrs(), cin, cin, wt, native.EPI_RELU, out_split=out.ptrs(), ldo_split=wt.coutpad, flags=fl)
        e1.record(); torch.cuda.synchronize()
        print(f"{cin}->{cout} {kh}x{kw}  split_n={fl >> 2}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us", flush=True)
