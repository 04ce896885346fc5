#!/usr/bin/env python
"""bench.py -- rendered faces/sec of the FENeRF volumetric render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model A|B]
                    [--precision guard|fast|exact]

A "step" is one pass of the hot path over one batch of synthetic latents: BASELINE.json configs[1]
-- 128x128 image, 24 (+24 hierarchical) samples per ray, batch 4 per GPU, forward-only -- through
the reference-facing generator API (``generator(z, **metadata)`` under no_grad), then the frame
all-gather when N > 1.  Weights are the reference's random init under manual_seed(0); latents are
N(0,1); camera poses gaussian (h_stddev 0.3, v_stddev 0.155); nerf_noise 0.  One JSON line on
stdout (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

FLOP_PER_POINT = {"A": 1053696, "B": 1341440}   # SURVEY.md section 8d (B: label chain pre-multiplied)
MODEL_NAME = {"A": "ImplicitGenerator3d+TALLSIREN", "B": "DoubleImplicitGenerator3d+TextureEmbeddingPiGAN256SEMANTICDISENTANGLE_DIM_96"}
IMG, STEPS_PER_RAY, BATCH_PER_GPU = 128, 24, 4


def metadata(img_size=IMG):
    return dict(img_size=img_size, fov=12, ray_start=0.88, ray_end=1.12, num_steps=STEPS_PER_RAY, h_stddev=0.3,
                v_stddev=0.155, h_mean=math.pi * 0.5, v_mean=math.pi * 0.5, hierarchical_sample=True,
                sample_dist='gaussian', clamp_mode='relu', nerf_noise=0.0, last_back=False)


def build_generator(model, device):
    from fenerf_b200.generators import generators as g
    from fenerf_b200.siren import siren as s
    torch.manual_seed(0)
    if model == "A":
        gen = g.ImplicitGenerator3d(s.TALLSIREN, 256, 4)
    else:
        gen = g.DoubleImplicitGenerator3d(s.TextureEmbeddingPiGAN256SEMANTICDISENTANGLE_DIM_96, 256, 256, 22)
    gen.eval()
    gen.to(device)
    gen.device = device
    gen.siren.device = device
    return gen


def make_latents(model, n_batches, batch, seed0=1000):
    out = []
    for i in range(n_batches):
        torch.manual_seed(seed0 + i)
        out.append(tuple(torch.randn(batch, 256) for _ in range(1 if model == "A" else 2)))
    return out


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path).read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        # the busy samples are the upper half (idle gaps between regions clock down)
        busy = sm[len(sm) // 2:]
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION and
        # =WARN, so drop those levels and send whatever NCCL logs at other levels to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def max_over_ranks(ms, device, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# ------------------------------------------------------------------------------------------------
def cpu_faces_per_sec(model, steps, warmup, budget_s=200.0):
    """The reference's CPU path (the oracle port: same ATen ops in the same order, all host
    threads) on a bounded sample of the cfg2 workload.  Returns (faces/s, description, cores)."""
    from oracle import render_oracle as oracle
    avail = os.cpu_count() or 1
    gen = build_generator(model, "cpu")
    lat = make_latents(model, 1, 1)[0]
    film = oracle.film_from_latents(gen.siren, lat)
    # Give the reference the thread count it runs best with: torch defaults to one thread per core,
    # which on a 100+-core host is slower than a smaller pool for these (P, 256) tensors.  Probe at
    # 32 px (per-ray cost is resolution independent), keep the fastest.
    probe, cores = None, avail
    for n in sorted({avail, 64, 32, 16, 8}, reverse=True):
        if n > avail:
            continue
        torch.set_num_threads(n)
        oracle.render(gen.siren, film, metadata(16))
        t0 = time.perf_counter()
        oracle.render(gen.siren, film, metadata(32))
        dt = time.perf_counter() - t0
        if probe is None or dt < probe:
            probe, cores = dt, n
    torch.set_num_threads(cores)
    per_face = probe * (IMG / 32) ** 2
    r = IMG
    while r > 32 and per_face * (r / IMG) ** 2 * (steps + warmup) > budget_s:
        r //= 2
    md = metadata(r)
    for _ in range(warmup):
        oracle.render(gen.siren, film, md)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.render(gen.siren, film, md)
    dt = (time.perf_counter() - t0) / steps
    frac = (r / IMG) ** 2
    sample = "%d step(s) of %dx%d rays (%.3g of one cfg2 face, B=1, %d+%d samples/ray), oracle port, fp32, %d threads (best of a thread-count probe on %d host cores)" % (
        steps, r, r, frac, STEPS_PER_RAY, STEPS_PER_RAY, cores, avail)
    return frac / dt, sample, cores


def run_reference_arm(args, world, rank):
    if rank != 0:
        return
    value, sample, cores = cpu_faces_per_sec(args.model, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "rendered faces/sec at 128px x 24 samples/ray", "value": value,
        "unit": "faces/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "faces/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {"workload": "cfg2: %s, %dx%d, %d+%d samples/ray hierarchical, batch %d/GPU, forward-only render" % (
                MODEL_NAME[args.model], IMG, IMG, STEPS_PER_RAY, STEPS_PER_RAY, BATCH_PER_GPU),
            "global_batch": BATCH_PER_GPU * world, "parallelism": "dp%d (images sharded, one frame all-gather)" % world,
            "precision_mode": args.precision,
            "l2_policy": "inputs_larger_than_l2 (per-step working set ~200 MB of samples/raw outputs vs 126 MB L2; "
                         "fresh latents and RNG draws every step)"}


# ------------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local):
    from fenerf_b200 import _lib, ops
    from fenerf_b200.dist import FrameGatherer
    from fenerf_b200.generators import volumetric_rendering as vr
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    _lib.lib()
    ops.set_default_precision(args.precision)
    gen = build_generator(args.model, device)
    md = metadata()
    B = BATCH_PER_GPU
    C_img = gen.output_dim - 1
    n_batches = args.steps + args.warmup
    lat_host = [tuple(z.pin_memory() for z in zs) for zs in make_latents(args.model, n_batches, B, 1000 + 97 * rank)]
    lat_dev = [tuple(z.to(device) for z in zs) for zs in lat_host]
    gatherer = FrameGatherer(B, C_img, IMG, device)
    out_host = [torch.empty((world * B, C_img, IMG, IMG), dtype=torch.float32).pin_memory() for _ in range(2)]
    out_done = [torch.cuda.Event() for _ in range(2)]
    frames_ready = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=device)
    host_sink = [0.0]
    torch.manual_seed(4242 + rank)

    def step_resident(i):
        with torch.no_grad():
            frames, _ = gen(*lat_dev[i], **md)
        return gatherer.gather(frames)

    def step_e2e(i):
        if i > first_e2e[0]:
            # the gathered-frame buffer is reused every step: wait (on the GPU) until the previous D2H has read it
            torch.cuda.current_stream().wait_event(out_done[(i - 1) & 1])
        zs = tuple(z.to(device, non_blocking=True) for z in lat_host[i])
        with torch.no_grad():
            frames, _ = gen(*zs, **md)
        allf = gatherer.gather(frames)
        frames_ready[i & 1].record()
        copy_stream.wait_event(frames_ready[i & 1])
        with torch.cuda.stream(copy_stream):        # D2H on its own stream: the next step's kernels do not queue behind it
            out_host[i & 1].copy_(allf, non_blocking=True)
            out_done[i & 1].record()
        # double-buffered serving loop: the host reads step i-1's frames while step i is queued, so the
        # GPU does not idle through the host's launch work; every step's frames reach the host and are read
        if i > first_e2e[0]:
            read_frames(i - 1)

    def read_frames(i):
        out_done[i & 1].synchronize()
        host_sink[0] += float(out_host[i & 1][0, 0, 0, 0]) + float(out_host[i & 1][-1, -1, -1, -1])

    first_e2e = [0]

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- arm 1: inputs resident in HBM ----
    for i in range(args.warmup):
        step_resident(i)
    barrier(world); torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_resident(args.warmup + i)
    ev1.record()
    torch.cuda.synchronize(); barrier(world)
    launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), device, world)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    # ---- arm 2: end to end through the public API with host buffers ----
    n_pre = min(args.warmup, 2)
    for i in range(n_pre):
        step_e2e(i)
    if n_pre:
        read_frames(n_pre - 1)
    barrier(world); torch.cuda.synchronize()
    first_e2e[0] = args.warmup
    ev0.record()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    read_frames(args.warmup + args.steps - 1)       # the last step's frames, inside the timed region
    ev1.record()
    torch.cuda.synchronize(); barrier(world)
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1), device, world) / args.steps
    e2e_value = world * B / (ms_e2e / 1e3)
    h2d = sum(z.numel() * 4 for z in lat_host[0])
    d2h = out_host[0].numel() * 4
    # ---- roofline of the dominant kernel: the point-network launches, CUDA events on their stream ----
    roof = field_roofline(gen, args, lat_dev[0], md, device)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    line = {
        "metric": "rendered faces/sec at 128px x 24 samples/ray", "value": value, "unit": "faces/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (tcgen05) + f32 refinement" if args.precision != "exact" else "f32",
        "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "faces/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "pipeline": "double-buffered pinned output, D2H on a copy stream: step i-1's frames are read on the host while step i runs"},
        "gpu_launches": int(launches), "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, sample, cores = cpu_faces_per_sec(args.model, 1, 0, budget_s=40.0)
        line["cpu_baseline"] = {"value": v, "unit": "faces/s", "cores": cores, "kind": "port", "sample": sample}
    emit(line)


def field_roofline(gen, args, latents, md, device):
    """Times the point-network launches of one step with CUDA events on the launching stream
    (same sizes and inputs as inside the step: ray_setup -> field -> resample -> field)."""
    from fenerf_b200 import ops
    from fenerf_b200.generators import volumetric_rendering as vr
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback (B200_PROFILING.md sustained ~1.4 PF)"
    B, S, R = latents[0].shape[0], md["num_steps"], md["img_size"]
    N = R * R
    with torch.no_grad():
        if args.model == "A":
            film = gen.siren.film_table(*gen.siren.mapping_network(latents[0]))
        else:
            fg, pg = gen.siren.geo_mapping_network(latents[0]); fa, pa = gen.siren.app_mapping_network(latents[1])
            film = gen.siren.film_table(fg, fa, pg, pa)
        rd = ops.make_render_desc(batch=B, img_size=R, num_steps=S, hierarchical=True, clamp_mode='relu', nerf_noise=0.0,
                                  fov=md["fov"], precision=args.precision)
        x_lin, y_lin, z_lin = vr.ray_tables(R, S, md["ray_start"], md["ray_end"], device)
        c2w, _, _ = ops.camera_poses(B, 'gaussian', 0.3, 0.155, md["h_mean"], md["v_mean"], vr.DeviceRng(device), device)
        durations = []
        for it in range(6):
            pts, z, dirs, org = ops.ray_setup(rd, x_lin, y_lin, z_lin, c2w, torch.rand(B, N, S, 1, device=device))
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            # keep the GPU busy (~2 ms spin) while the host queues the launches below, so the events
            # bracket kernel execution and not Python launch latency on an idle GPU
            torch.cuda._sleep(4_000_000)
            e[0].record()
            raw_c = ops.siren_points(gen.siren, pts.reshape(B, N * S, 3), film, dirs, precision=rd.precision)
            e[1].record()
            z_f, pts_f, _ = ops.resample(rd, raw_c.reshape(B, N, S, -1), z, dirs, org, None, torch.rand(B * N, S, device=device))
            e[2].record()
            ops.siren_points(gen.siren, pts_f.reshape(B, N * S, 3), film, dirs, precision=rd.precision)
            e[3].record()
            torch.cuda.synchronize()
            if it > 0:
                durations += [e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])]
    durations.sort()
    ms = durations[len(durations) // 2] if len(durations) % 2 else 0.5 * (durations[len(durations) // 2 - 1] + durations[len(durations) // 2])
    flops = B * N * S * FLOP_PER_POINT[args.model]
    achieved = flops / (ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s_%s" % (args.model, args.precision))
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "siren point network (%s)" % args.precision, "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "ms_per_launch": ms,
            "flop_per_launch": flops, "peak_source": peak_src,
            "frac_of_burst_peak": achieved / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None}


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries underneath write there too (NCCL's version banner,
    whatever the box's NCCL_DEBUG is), so the real stdout is kept aside and fd 1 points at stderr until
    `emit` writes the line."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="A", choices=["A", "B"])
    ap.add_argument("--precision", default=os.environ.get("FENERF_B200_PRECISION", "guard"), choices=["guard", "fast", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        run_reference_arm(args, world, rank)
        return
    world, rank, local = dist_setup(args.gpus)
    try:
        run_ours(args, world, rank, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
