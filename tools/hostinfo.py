import os, time, sys, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
os.system("grep -m1 'model name' /proc/cpuinfo; free -g | head -2")
sys.path.insert(0, '.'); sys.path.insert(0, 'raft-ncup_b200'); sys.path.insert(0, 'tests')
from rnc.synth import build_model, frames
from oracle import raft_oracle as orc
sd = {k: v.detach() for k, v in build_model("raft_nc_dbl").state_dict().items()}
im1, im2 = frames(1, 128, 256)
for nt in (None, 8, 16, 32, 64):
    if nt: torch.set_num_threads(nt)
    t=time.perf_counter(); orc.raft_forward(sd, im1, im2, iters=4); print("threads", torch.get_num_threads(), "cfg1 4 iters", round(time.perf_counter()-t,3), flush=True)
