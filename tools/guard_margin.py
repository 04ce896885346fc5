"""How far is the tcgen05 path's density from the fp32 path's?  The GUARD mode re-evaluates a ray's far sample
in fp32 when |sigma_far| < tau (default 1.5e-3); this prints the error distribution that tau has to cover,
over points sampled like render points (inside the camera frustum volume) for every field class."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cases
from fenerf_b200 import ops

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    torch.manual_seed(7)
    for name in ("a_small", "b_small", "c_small", "d_small"):
        case = _cases.CASE_BY_NAME[name]
        gen = _cases.build_mirror(case, "cuda:0")
        worst = 0.0
        qs = []
        for trial in range(4):                                   # four latent codes
            lat = [torch.randn(1, 256, device="cuda") for _ in range(_cases.n_latents(case.model))]
            pts = (torch.rand(1, n // 4, 3, device="cuda") - 0.5) * 0.24
            with torch.no_grad():
                if len(lat) == 1:
                    film = gen.siren.film_table(*gen.siren.mapping_network(lat[0]))
                else:
                    fg, pg = gen.siren.geo_mapping_network(lat[0]); fa, pa = gen.siren.app_mapping_network(lat[1])
                    film = gen.siren.film_table(fg, fa, pg, pa)
                fast = ops.siren_sigma(gen.siren, pts, film, precision="fast")
                exact = ops.siren_sigma(gen.siren, pts, film, precision="exact")
            err = (fast - exact).abs().flatten()
            worst = max(worst, err.max().item())
            qs.append(torch.quantile(err[:: max(1, err.numel() // 1_000_000)], torch.tensor([0.5, 0.99, 0.9999], device="cuda")).cpu())
            flips = ((fast.flatten() > 0) != (exact.flatten() > 0))
            band = exact.flatten().abs()[flips]
            worst_flip = band.max().item() if flips.any() else 0.0
        q = torch.stack(qs).mean(0)
        print("%-8s %-48s n=%d  |d sigma|: median %.2e  p99 %.2e  p99.99 %.2e  max %.2e   largest |sigma_exact| among sign flips %.2e  (tau 1.5e-3)" % (
            name, type(gen.siren).__name__, n, q[0], q[1], q[2], worst, worst_flip))

main()
