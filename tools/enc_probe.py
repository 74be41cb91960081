"""Time the (cuDNN, strict fp32) encoders under a few backend settings.  Scratch tool for DESIGN.md's encoder notes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "raft-ncup_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from rnc.synth import build_model  # noqa: E402

dev = "cuda:0"
m = build_model("raft_nc_dbl").to(dev)
x1 = torch.rand(8, 3, 440, 1024, device=dev) * 2 - 1
x2 = torch.rand(8, 3, 440, 1024, device=dev) * 2 - 1


def run(label, bench, cl, tf32=False):
    torch.backends.cudnn.benchmark = bench
    a, b = (x1, x2) if not cl else (x1.contiguous(memory_format=torch.channels_last), x2.contiguous(memory_format=torch.channels_last))
    net = m if not cl else m.to(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3):
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32, benchmark=bench):
                f = net.fnet.conv2(net.fnet.layer3(net.fnet.layer2(net.fnet.layer1(net.fnet.relu1(net.fnet.norm1(net.fnet.conv1(torch.cat([a, b], 0))))))))
                c = net.cnet.conv2(net.cnet.layer3(net.cnet.layer2(net.cnet.layer1(net.cnet.relu1(net.cnet.norm1(net.cnet.conv1(a)))))))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32, benchmark=bench):
                f = net.fnet.conv2(net.fnet.layer3(net.fnet.layer2(net.fnet.layer1(net.fnet.relu1(net.fnet.norm1(net.fnet.conv1(torch.cat([a, b], 0))))))))
                c = net.cnet.conv2(net.cnet.layer3(net.cnet.layer2(net.cnet.layer1(net.cnet.relu1(net.cnet.norm1(net.cnet.conv1(a)))))))
        torch.cuda.synchronize()
    print(f"{label:40s} {(time.perf_counter() - t0) / 5 * 1e3:8.2f} ms", flush=True)
    return f, c


f0, c0 = run("fp32 default", False, False)
f1, c1 = run("fp32 cudnn.benchmark", True, False)
print("   diff vs default:", (f1 - f0).abs().max().item(), (c1 - c0).abs().max().item())
f2, c2 = run("fp32 benchmark + channels_last", True, True)
print("   diff vs default:", (f2 - f0).abs().max().item(), (c2 - c0).abs().max().item())
f3, c3 = run("TF32 benchmark + channels_last (ref only)", True, True, tf32=True)
print("   diff vs default:", (f3 - f0).abs().max().item(), (c3 - c0).abs().max().item())
