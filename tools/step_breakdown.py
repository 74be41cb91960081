"""Where does one benchmark step go?  CUDA-event brackets around the encoders, every corr lookup, every update-block
iteration and the NCUP upsampler (developer tool; the brackets are the Engine.profile hooks bench.py uses)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from rnc.synth import build_model, frames

B, H, W, iters = int(os.environ.get("B", 8)), 440, 1024, 32
dev = "cuda:0"
m = build_model("raft_nc_dbl").to(dev)
im1, im2 = frames(B, H, W)
im1, im2 = im1.to(dev), im2.to(dev)
eng = m.engine()
with torch.no_grad():
    for _ in range(3):
        m(im1, im2, iters=iters, test_mode=True)
    torch.cuda.synchronize()
    eng.profile = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m(im1, im2, iters=iters, test_mode=True)
    e1.record()
    torch.cuda.synchronize()
total = e0.elapsed_time(e1)
print(f"step {total:.2f} ms (B={B}, {iters} iters, brackets add a little)")
acc = 0.0
for name, evs in eng.profile.items():
    t = sum(a.elapsed_time(b) for a, b in evs)
    acc += t
    print(f"  {name:14s} x{len(evs):3d}  {t:8.3f} ms  ({100 * t / total:4.1f}%)  avg {t / len(evs) * 1e3:8.1f} us")
print(f"  {'unbracketed':14s}       {total - acc:8.3f} ms")
