"""Host enqueue time of one eager forward (no graph) against its GPU time: how much slack does the launch loop have?"""
import os, sys, time, torch
os.environ.setdefault("RNC_GRAPH", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200")):
    sys.path.insert(0, p)
from rnc.synth import build_model, frames
B = int(os.environ.get("B", 8))
m = build_model("raft_nc_dbl").to("cuda:0")
im1, im2 = frames(B, 440, 1024)
im1, im2 = im1.to("cuda:0"), im2.to("cuda:0")
with torch.no_grad():
    for _ in range(3):
        m(im1, im2, iters=32, test_mode=True)
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        m(im1, im2, iters=32, test_mode=True)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"B={B}: enqueue {1e3 * (t1 - t0):.2f} ms, until done {1e3 * (t2 - t0):.2f} ms", flush=True)
