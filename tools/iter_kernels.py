"""Warm per-layer times of the update block: CUDA-event brackets around every rnc_conv2d_umma_fwd call of the iterations
(developer tool; the brackets cost ~2 us each and defeat PDL overlap, so the sum exceeds the real iteration time)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from rnc.synth import build_model, frames

B, iters = int(os.environ.get("B", 8)), 12
m = build_model("raft_nc_dbl").to("cuda:0")
im1, im2 = frames(B, 440, 1024)
im1, im2 = im1.to("cuda:0"), im2.to("cuda:0")
eng = m.engine()
rec = []
orig = eng.uconv
def timed(B_, H, W, in0, c0, ld0, wt, epi, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(B_, H, W, in0, c0, ld0, wt, epi, **kw); e1.record()
    rec.append((f"{c0 + kw.get('c1', 0)}->{wt.cout} {wt.kh}x{wt.kw} epi{epi} {H}x{W}", e0, e1))
with torch.no_grad():
    m(im1, im2, iters=iters, test_mode=True)
    torch.cuda.synchronize()
    eng.uconv = timed
    m(im1, im2, iters=iters, test_mode=True)
    torch.cuda.synchronize()
names = [r[0] for r in rec]
# the loop's layers repeat with period = calls per iteration: find the tail period
loop = [r for r in rec if r[0].endswith("55x128")]
first = [i for i, r in enumerate(loop) if r[0] == loop[-1][0]]
per = first[-1] - first[-2]                     # calls per iteration = distance between repeats of the last layer
body = loop[-per * (iters - 2):]
tot = 0.0
for k in range(per):
    ts = [body[i][1].elapsed_time(body[i][2]) * 1e3 for i in range(k, len(body), per)]
    tot += sum(ts) / len(ts)
    print(f"{body[k][0]:32s} {sum(ts) / len(ts):7.1f} us")
print(f"sum {tot:.1f} us per iteration (bracketed)")
