"""Where does a training step spend its GPU time?  torch.profiler (CUPTI) over one cfg-5 step: kernel totals."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "raft-ncup_b200")]
from rnc.synth import build_model  # noqa: E402
from rnc.train import fetch_optimizer, train_step  # noqa: E402

dev = torch.device("cuda", 0)
m = build_model("raft_nc_dbl").to(dev).train()
m.freeze_bn()
opt, sched = fetch_optimizer(m, lr=1e-4, num_steps=100)
g = torch.Generator().manual_seed(1)
B, H, W = 2, 384, 512
im1, im2 = (torch.rand(B, 3, H, W, generator=g) * 255).to(dev), (torch.rand(B, 3, H, W, generator=g) * 255).to(dev)
gt, valid = (torch.randn(B, 2, H, W, generator=g) * 5).to(dev), torch.ones(B, H, W, device=dev)
for _ in range(3):
    train_step(m, opt, sched, im1, im2, gt, valid, iters=12, return_metrics=False)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    train_step(m, opt, sched, im1, im2, gt, valid, iters=12, return_metrics=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
