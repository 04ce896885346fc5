"""Per-role clock64 trace of CTA 0 of the tcgen05 point-network kernel (fenerf_debug_trace)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cases
from fenerf_b200 import _lib, ops
model = sys.argv[1] if len(sys.argv) > 1 else "A"
n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 0
show_tile = int(sys.argv[3]) if len(sys.argv) > 3 else 1
case = _cases.CASE_BY_NAME["a_small" if model == "A" else "b_small"]
gen = _cases.build_mirror(case, "cuda:0")
B, N, S = 4, 128 * 128, 24
if n_tiles:
    B, N, S = 1, n_tiles * 128 // 8, 8
pts = (torch.rand(B, N * S, 3, device="cuda") - 0.5) * 0.3
dirs = torch.nn.functional.normalize(torch.randn(B, N, 3, device="cuda"), dim=-1)
with torch.no_grad():
    if model == "A":
        film = gen.siren.film_table(*gen.siren.mapping_network(torch.randn(B, 256, device="cuda")))
    else:
        fg, pg = gen.siren.geo_mapping_network(torch.randn(B, 256, device="cuda")); fa, pa = gen.siren.app_mapping_network(torch.randn(B, 256, device="cuda"))
        film = gen.siren.film_table(fg, fa, pg, pa)
    for _ in range(2):
        ops.siren_points(gen.siren, pts, film, dirs, precision="fast")
    torch.cuda.synchronize()
    buf = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
    _lib.lib().fenerf_debug_trace(buf.data_ptr())
    ops.siren_points(gen.siren, pts, film, dirs, precision="fast")
    torch.cuda.synchronize()
    _lib.lib().fenerf_debug_trace(0)
t = buf.cpu().reshape(4, 4096)
ev = []
for role in range(4):
    n = int(t[role, 0])
    for i in range(n):
        tag, clk = int(t[role, 2 + 2 * i]), int(t[role, 3 + 2 * i])
        ev.append((clk, role, chr(tag >> 48), (tag >> 32) & 0xffff, (tag >> 16) & 0xffff, tag & 0xffff))
ev.sort()
t0 = ev[0][0]
names = ["mmaY", "mmaX", "epiX", "epiY"]
for clk, role, kind, tile, stage, item in ev:
    if tile == show_tile:
        print("%9d %s %s tile%d stage%2d item%3d" % (clk - t0, names[role], kind, tile, stage, item))
