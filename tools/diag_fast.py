"""GPU diagnostic: tcgen05 (fast) vs CUDA-core (exact) point network on the same inputs, plus timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cases, _harness
from fenerf_b200 import ops

def main():
    names = sys.argv[1:] or ["a_small", "b_small"]
    for name in names:
        case = _cases.CASE_BY_NAME[name]
        run = _harness.oracle_run(case)
        st = run["out"]["stages"]
        gen = _cases.build_mirror(case, "cuda:0")
        b, n, s = case.batch, case.cfg["img_size"] ** 2, case.cfg["num_steps"]
        pts = st["points_coarse"].reshape(b, n * s, 3).contiguous().cuda()
        film = run["film"].cuda(); dirs = st["dirs"].contiguous().cuda()
        with torch.no_grad():
            ex = ops.siren_points(gen.siren, pts, film, dirs, precision="exact")
            torch.cuda.synchronize()
            fa = ops.siren_points(gen.siren, pts, film, dirs, precision="fast")
            torch.cuda.synchronize()
        ref = st["raw_coarse"].reshape(b, n * s, -1)
        e_ex = (ex.cpu() - ref).abs(); e_fa = (fa.cpu() - ref).abs()
        print("[%s] exact vs oracle: max %.3e | fast vs oracle: max %.3e mean %.3e" % (name, e_ex.max(), e_fa.max(), e_fa.mean()))
        print("   per-channel max (fast):", ["%.2e" % v for v in e_fa.amax((0, 1)).tolist()])
        print("   sample fast:", fa[0, :2].cpu().tolist()); print("   sample ref :", ref[0, :2].tolist())
    # throughput of the field kernels alone at cfg2 size (model A)
    case = _cases.CASE_BY_NAME["a_small"]
    gen = _cases.build_mirror(case, "cuda:0")
    B, N, S = 4, 128 * 128, 24
    pts = (torch.rand(B, N * S, 3, device="cuda") - 0.5) * 0.3
    dirs = torch.nn.functional.normalize(torch.randn(B, N, 3, device="cuda"), dim=-1)
    z = torch.randn(B, 256, device="cuda")
    with torch.no_grad():
        film = gen.siren.film_table(*gen.siren.mapping_network(z))
        for prec in ("fast", "exact"):
            for _ in range(2):
                ops.siren_points(gen.siren, pts, film, dirs, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if prec == "fast" else 2
            e0.record()
            for _ in range(reps):
                ops.siren_points(gen.siren, pts, film, dirs, precision=prec)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = B * N * S * 1053696 / (ms * 1e-3) / 1e12
            print("field[%s] %d pts: %.3f ms  %.1f TFLOP/s" % (prec, B * N * S, ms, tf))

if __name__ == "__main__":
    main()
