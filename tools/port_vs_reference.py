"""Build container only: times the oracle port (bench.py's cpu_baseline / --impl reference arm) against the LIVE, unmodified
reference on the same host and threads, so that the port's cost can be trusted as the reference's.
    python tools/port_vs_reference.py > profiles/r02_port_vs_reference.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cases
from oracle import ref_shim, render_oracle as oracle

ref_generators, ref_siren, _ = ref_shim.load()
threads = min(os.cpu_count() or 1, 8)
torch.set_num_threads(threads)
print("host: %d threads, torch %s" % (threads, torch.__version__))
for model, img, steps in (("A", 64, 12), ("A", 128, 24), ("B", 64, 12)):
    name = {"A": "a_small", "B": "b_small"}[model]
    case = _cases.CASE_BY_NAME[name]
    torch.manual_seed(0)
    gen_ref = _cases.construct(ref_generators, ref_siren, model)
    gen_ref.set_device("cpu"); gen_ref.eval()
    gen = _cases.build_mirror(case, "cpu")
    cfg = dict(case.cfg, img_size=img, num_steps=steps)
    lat = tuple(z[:1] for z in _cases.make_latents(case))
    film = oracle.film_from_latents(gen.siren, lat)

    def t_ref():
        with torch.no_grad():
            gen_ref(*lat, **cfg)

    def t_port():
        oracle.render(gen.siren, film, cfg)

    res = {}
    for label, fn in (("reference", t_ref), ("oracle port", t_port)):
        fn()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
        res[label] = best
    print("model %s %3dx%-3d %2d+%-2d samples, B=1:  reference %.3f s   oracle port %.3f s   port/reference = %.3f" % (
        model, img, img, steps, steps, res["reference"], res["oracle port"], res["oracle port"] / res["reference"]))
