"""Where the end-to-end arm's extra time per step goes: bench.py's e2e loop with parts switched off, and the host's time split
into queueing vs. blocked-on-the-GPU.   python tools/diag_e2e.py [A|B] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import bench

model = sys.argv[1] if len(sys.argv) > 1 else "A"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
args = argparse.Namespace(precision="guard", model=model, no_graph=False)
dev = torch.device("cuda", 0)
r = bench.StepRunner(args, model, 1, 0, dev, 64, use_graph=True)


def loop(h2d=True, d2h=True, stage=True, read_lag=1, n=steps):
    """One e2e-shaped loop; returns (ms/step on the device, host ms/step queueing, host ms/step blocked)."""
    lat = r.lat_host if h2d else r.lat_dev
    t_queue = t_block = 0.0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        t0 = time.perf_counter()
        frames = r.graph(*lat[i % len(lat)])[0]
        if stage:
            r.stage[i & 1].copy_(frames)
        if d2h:
            r.frames_ready[i & 1].record()
            r.copy_stream.wait_event(r.frames_ready[i & 1])
            with torch.cuda.stream(r.copy_stream):
                r.out_host[i & 1].copy_(r.stage[i & 1] if stage else frames, non_blocking=True)
                r.out_done[i & 1].record()
        t1 = time.perf_counter()
        if d2h and i >= read_lag and read_lag < 2:
            r.out_done[(i - read_lag) & 1].synchronize()
            r.host_sink += float(r.out_host[(i - read_lag) & 1][0, 0, 0, 0])
        t2 = time.perf_counter()
        t_queue += t1 - t0
        t_block += t2 - t1
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * t_queue / n, 1e3 * t_block / n


def resident(n=steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        r.graph(*r.lat_dev[i % len(r.lat_dev)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(5):
    resident(5)
rows = []
for rep in range(2):
    rows.append(("resident (graph replay only)", (resident(), 0.0, 0.0)))
    rows.append(("e2e as in bench (read lag 1)", loop()))
    rows.append(("e2e, host never reads inside the loop", loop(read_lag=9)))
    rows.append(("e2e, latents resident (no H2D)", loop(h2d=False)))
    rows.append(("e2e, no D2H / no read", loop(d2h=False)))
    rows.append(("e2e, no staging copy", loop(stage=False)))
    rows.append(("H2D + graph only", loop(d2h=False, stage=False)))
for name, (ms, q, b) in rows:
    print("%-42s %.4f ms/step on the device   host: queue %.3f ms  blocked %.3f ms" % (name, ms, q, b))
