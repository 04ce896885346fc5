"""A/B of the backward's 256-wide products inside one process: the library's tcgen05 GEMMs (csrc/gemm5.cu) against cuBLAS
(torch.mm), alternating, on the differentiable render of bench.py's cfg3-shaped training step (model B, 64x64, 24+24, batch 8).
Times forward+backward of one split with CUDA events; the host queues ahead, so this is GPU time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from fenerf_b200 import backward

dev = torch.device("cuda:0")
gen = bench.build_generator("B", dev)
md = dict(bench.metadata(64), num_steps=24)
lat = [z.to(dev) for z in bench.make_latents("B", 1, 8)[0]]

def step():
    for p in gen.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.float16):
        frames, _ = gen(*lat, **md)
        loss = (frames.float() ** 2).mean()
    loss.backward()

def run(mode, reps=8):
    backward.BWD_GEMM = mode
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {"tcgen05": [], "cublas": []}
for rnd in range(4):
    for mode in ("tcgen05", "cublas"):
        res[mode].append(run(mode))
for mode, v in res.items():
    print("%-8s forward+backward of one split (8 faces, 64x64, 24+24): %s ms   median %.2f" % (mode, " ".join("%.2f" % x for x in v), sorted(v)[len(v) // 2]))
