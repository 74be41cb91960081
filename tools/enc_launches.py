"""One forward with iters=1 (encoders + one update iteration + upsampler), twice; run under
ncu --metrics gpu__time_duration.sum to list the encoder-phase launches (developer tool)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from rnc.synth import build_model, frames
B = int(os.environ.get("B", 8))
m = build_model("raft_nc_dbl").to("cuda:0")
im1, im2 = frames(B, 440, 1024)
im1, im2 = im1.to("cuda:0"), im2.to("cuda:0")
with torch.no_grad():
    for _ in range(2):
        m(im1, im2, iters=1, test_mode=True)
        torch.cuda.synchronize()
