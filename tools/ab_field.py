"""Times the tcgen05 point-network launch alone (cfg2 sizes), median of N, for A/B experiments:
    FENERF_B200_LIB=fenerf_b200/lib_x.so python tools/ab_field.py [A|B] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from fenerf_b200 import ops
from fenerf_b200.generators import volumetric_rendering as vr

model = sys.argv[1] if len(sys.argv) > 1 else "A"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
gen = bench.build_generator(model, dev)
md = bench.metadata()
B, S, R = 4, md["num_steps"], md["img_size"]
N = R * R
lat = [z.to(dev) for z in bench.make_latents(model, 1, B)[0]]
with torch.no_grad():
    if model == "A":
        film = gen.siren.film_table(*gen.siren.mapping_network(lat[0]))
    else:
        fg, pg = gen.siren.geo_mapping_network(lat[0]); fa, pa = gen.siren.app_mapping_network(lat[1])
        film = gen.siren.film_table(fg, fa, pg, pa)
    rd = ops.make_render_desc(batch=B, img_size=R, num_steps=S, hierarchical=True, clamp_mode='relu', nerf_noise=0.0, fov=12, precision="fast")
    x, y, z = vr.ray_tables(R, S, md["ray_start"], md["ray_end"], dev)
    c2w, _, _ = ops.camera_poses(B, 'gaussian', 0.3, 0.155, md["h_mean"], md["v_mean"], vr.DeviceRng(dev), dev)
    pts, zv, dirs, org = ops.ray_setup(rd, x, y, z, c2w, torch.rand(B, N, S, 1, device=dev))
    ts = []
    for it in range(n + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)
        e0.record()
        out = ops.siren_points(gen.siren, pts.reshape(B, N * S, 3), film, dirs, precision="fast")
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
ts.sort()
flops = B * N * S * bench.FLOP_PER_POINT[model]
print("%s model %s: median %.4f ms  min %.4f  max %.4f  -> %.1f TFLOP/s   checksum %.6f" % (
    os.environ.get("FENERF_B200_LIB", "default"), model, ts[len(ts) // 2], ts[0], ts[-1], flops / ts[len(ts) // 2] / 1e9, float(out.double().sum())))
