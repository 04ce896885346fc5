"""Writes the tcgen05 point-network outputs of a fixed input to a file (for comparing two builds):
    FENERF_B200_LIB=... python tools/dump_field.py out.pt [n_points]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cases
from fenerf_b200 import ops
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
case = _cases.CASE_BY_NAME["a_small"]
gen = _cases.build_mirror(case, "cuda:0")
g = torch.Generator(device="cuda").manual_seed(3)
pts = (torch.rand(1, n, 3, device="cuda", generator=g) - 0.5) * 0.3
dirs = torch.nn.functional.normalize(torch.randn(1, n, 3, device="cuda", generator=g), dim=-1)
z = torch.randn(1, 256, device="cuda", generator=g)
with torch.no_grad():
    film = gen.siren.film_table(*gen.siren.mapping_network(z))
    out = ops.siren_points(gen.siren, pts, film, dirs, precision="fast")
    ref = ops.siren_points(gen.siren, pts, film, dirs, precision="exact")
torch.cuda.synchronize()
torch.save({"out": out.cpu(), "ref": ref.cpu()}, sys.argv[1])
err = (out - ref).abs().amax(-1)[0].cpu()
bad = (err > 5e-3) | torch.isnan(err)
print("n", n, "bad points", int(bad.sum()), "of", n, "max err", float(torch.nan_to_num(err, nan=9.0).max()))
for t0 in range(0, n, 128):
    e = err[t0:t0 + 128]
    print("tile %d: bad %d  first bad rows %s" % (t0 // 128, int(bad[t0:t0 + 128].sum()), bad[t0:t0 + 128].nonzero().flatten()[:10].tolist()))
