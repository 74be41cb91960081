"""Does capturing the whole forward in a CUDA graph shorten the step?  (launch-gap probe)"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from rnc.synth import build_model
sys.argv = [sys.argv[0]]
import bench
dev = "cuda:0"
m = build_model("raft_nc_dbl").to(dev)
p1, p2 = bench.synth_frames(8, 7)
d1, d2 = p1.to(dev), p2.to(dev)
def run():
    with torch.no_grad():
        return m(d1, d2, iters=32, test_mode=True)
for _ in range(3): run()
torch.cuda.synchronize()
def timeit(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3
print("eager  : %.2f ms (gpu events)  %.2f ms (wall)" % timeit(run))
# host-only cost of issuing one forward (no sync inside): time until the python call returns
torch.cuda.synchronize(); t0 = time.perf_counter(); run(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("host issue time of one forward: %.2f ms" % ((t1 - t0) * 1e3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    run()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = run()
torch.cuda.synchronize()
print("graph  : %.2f ms (gpu events)  %.2f ms (wall)" % timeit(g.replay))
ref = run()
g.replay(); torch.cuda.synchronize()
print("graph == eager:", (out[1] - ref[1]).abs().max().item())
