"""Kernel-time breakdown of one cfg3-shaped training iteration (torch.profiler, CUDA activities)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

dev = torch.device("cuda:0")
gen = bench.build_generator("B", dev); gen.train()
R, BATCH, SPLIT = 64, 32, 4
md = dict(bench.metadata(R), precision="guard")
opt = torch.optim.Adam(gen.parameters(), lr=6e-5, betas=(0.0, 0.9))
scaler = torch.amp.GradScaler("cuda")
w = torch.randn((BATCH // SPLIT, gen.output_dim - 1, R, R), device=dev) / (R * R)

def iteration(parts=("nograd", "grad", "opt")):
    if "nograd" in parts:
        for _ in range(2):
            with torch.no_grad():
                for _ in range(SPLIT):
                    gen(torch.randn(8, 256, device=dev), torch.randn(8, 256, device=dev), **md)
    if "grad" in parts:
        opt.zero_grad(set_to_none=True)
        for _ in range(SPLIT):
            with torch.autocast("cuda", dtype=torch.float16):
                px, _ = gen(torch.randn(8, 256, device=dev), torch.randn(8, 256, device=dev), **md)
                loss = (px * w).sum()
            scaler.scale(loss).backward()
    if "opt" in parts:
        scaler.unscale_(opt); torch.nn.utils.clip_grad_norm_(gen.parameters(), 10); scaler.step(opt); scaler.update()

for _ in range(2):
    iteration()
torch.cuda.synchronize()
for parts in (("nograd",), ("grad",), ("opt",)):
    if parts == ("opt",):
        iteration(("grad",))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); iteration(parts); e1.record(); torch.cuda.synchronize()
    print("part %-8s %8.2f ms" % (parts[0], e0.elapsed_time(e1)))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    iteration(("grad",))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
