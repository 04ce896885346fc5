"""In-step (warm) stage durations of fenerf_render_forward, CUDA events between the launches (fenerf_debug_stage_times):
    python tools/stage_times.py [A|B] [img] [steps_per_ray] [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from fenerf_b200 import _lib

model = sys.argv[1] if len(sys.argv) > 1 else "A"
img = int(sys.argv[2]) if len(sys.argv) > 2 else 128
spr = int(sys.argv[3]) if len(sys.argv) > 3 else 24
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
gen = bench.build_generator(model, dev)
md = dict(bench.metadata(img), num_steps=spr)
lat = [z.to(dev) for z in bench.make_latents(model, 1, B)[0]]
lib = _lib.lib()
names = ["ray_setup", "field_coarse", "guard", "resample", "field_fine", "composite"]
acc = [[] for _ in names]
tot = []
with torch.no_grad():
    for _ in range(3):
        gen(*lat, **md)
    lib.fenerf_debug_stage_times(1, None)
    for it in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(3_000_000)          # keep the GPU busy while the host queues the whole step
        e0.record()
        gen(*lat, **md)
        e1.record()
        out = (ctypes.c_float * 6)()
        lib.fenerf_debug_stage_times(1, out)
        torch.cuda.synchronize()
        for i in range(6):
            acc[i].append(out[i])
        tot.append(e0.elapsed_time(e1))
    lib.fenerf_debug_stage_times(0, None)
med = lambda v: sorted(v)[len(v) // 2]
print("model %s  %dx%d  %d+%d samples  batch %d   (median of 20, ms)" % (model, img, img, spr, spr, B))
for n, v in zip(names, acc):
    print("  %-14s %8.4f" % (n, med(v)))
print("  %-14s %8.4f   (sum of stages %.4f; the rest is RNG draws + mapping network + camera)" % ("whole call", med(tot), sum(med(v) for v in acc)))
