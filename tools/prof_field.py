"""One tcgen05 point-network launch at cfg2 size (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cases
from fenerf_b200 import ops
model = sys.argv[1] if len(sys.argv) > 1 else "A"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = _cases.CASE_BY_NAME["a_small" if model == "A" else "b_small"]
gen = _cases.build_mirror(case, "cuda:0")
B, N, S = 4, 128 * 128, 24
pts = (torch.rand(B, N * S, 3, device="cuda") - 0.5) * 0.3
dirs = torch.nn.functional.normalize(torch.randn(B, N, 3, device="cuda"), dim=-1)
with torch.no_grad():
    if model == "A":
        film = gen.siren.film_table(*gen.siren.mapping_network(torch.randn(B, 256, device="cuda")))
    else:
        fg, pg = gen.siren.geo_mapping_network(torch.randn(B, 256, device="cuda"))
        fa, pa = gen.siren.app_mapping_network(torch.randn(B, 256, device="cuda"))
        film = gen.siren.film_table(fg, fa, pg, pa)
    for _ in range(reps):
        ops.siren_points(gen.siren, pts, film, dirs, precision="fast")
torch.cuda.synchronize()
print("done")
