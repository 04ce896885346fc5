"""Runs the oracle for a parity case (CPU) -- shared by the CPU and the GPU tests."""
import hashlib
import random

import torch

import _cases
from oracle import render_oracle as oracle


def state_digest(module):
    h = hashlib.sha256()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def oracle_run(case, gen=None, keep_stages=True):
    """-> dict(film, latents, out, draws, avg_draws). Everything on the CPU, fp32."""
    gen = gen or _cases.build_mirror(case, "cpu")
    latents = _cases.make_latents(case)
    torch.manual_seed(case.seed)
    random.seed(case.seed)          # 'hybrid' camera mode flips Python's global coin (volumetric_rendering.py:199)
    avg_draws = None
    with torch.no_grad():
        film = oracle.film_from_latents(gen.siren, latents)
        if case.method == "staged_forward" and _cases.has_avg_frequencies(case):
            # generate_avg_frequencies draws first, then psi-truncation towards the mean FiLM
            # parameters (generators.py:142-149, 554-564); equivalent in table form because
            # 15 (a + psi (f - a)) + 30 is evaluated from the truncated raw frequency
            avg_draws = _cases.avg_film_draws(case)
            if _cases.n_latents(case.model) == 1:
                f, p = gen.siren.mapping_network(latents[0])
                fa, pa = gen.siren.mapping_network(avg_draws[0])
                fa, pa = fa.mean(0, keepdim=True), pa.mean(0, keepdim=True)
                f, p = fa + case.psi * (f - fa), pa + case.psi * (p - pa)
                film = gen.siren.film_table(f, p)
            else:
                fg, pg = gen.siren.geo_mapping_network(latents[0])
                fp, pp = gen.siren.app_mapping_network(latents[1])
                ag, bg = gen.siren.geo_mapping_network(avg_draws[0])
                ap, bp = gen.siren.app_mapping_network(avg_draws[1])
                ag, bg, ap, bp = (t.mean(0, keepdim=True) for t in (ag, bg, ap, bp))
                fg, pg = ag + case.psi * (fg - ag), bg + case.psi * (pg - bg)
                fp, pp = ap + case.psi * (fp - ap), bp + case.psi * (pp - bp)
                film = gen.siren.film_table(fg, fp, pg, pp)
    cfg = dict(case.cfg)
    cfg["softmax_label"] = bool(case.cfg.get("softmax_label", False))
    if case.method != "staged_forward":
        cfg.pop("fill_mode", None)
    out = oracle.render(gen.siren, film, cfg, keep_stages=keep_stages)
    return dict(gen=gen, film=film, latents=latents, out=out, draws=out["draws"], avg_draws=avg_draws)
