"""Where does the tcgen05 convolution's residual error come from?  Compare (a) full hi/lo split operands with
(b) operands that are exactly representable in fp16 (lo == 0: every product is exact, only the accumulation can err)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from test_gpu_umma_modes import ref_conv, run_conv  # noqa: E402
from rnc.engine_umma import UmmaEngine  # noqa: E402

eng = UmmaEngine()
g = torch.Generator().manual_seed(0)
for (cin, cout, kh, kw) in ((256, 192, 3, 3), (384, 256, 1, 5), (64, 64, 3, 3), (324, 256, 1, 1)):
    x = torch.randn(2, cin, 24, 128, generator=g)
    w = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.zeros(cout)
    ref = ref_conv(x, w, b)
    out, _ = run_conv(eng, x, w, b)
    xr, wr = x.half().float(), w.half().float()
    # weights are pre-scaled by a power of two inside UmmaWeights, so fp16-exact w stays exact
    ref_r = ref_conv(xr, wr, b)
    out_r, _ = run_conv(eng, xr, wr, b)
    d, dr = (out - ref), (out_r - ref_r)
    print(f"{cin}->{cout} {kh}x{kw}: split max {d.abs().max():.2e} mean {d.mean():+.2e} | fp16-exact inputs max {dr.abs().max():.2e} "
          f"mean {dr.mean():+.2e} | corr(sign(ref), err) {torch.sign(ref).mul(dr).mean():+.2e}  (|ref| mean {ref.abs().mean():.2f})")
