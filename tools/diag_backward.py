"""GPU diagnostic: the two halves of the backward against torch autograd in fp64 (composite) / fp32 (field).
    python tools/diag_backward.py [case ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import _cases, _harness  # noqa: E402
from fenerf_b200 import _lib, ops, backward  # noqa: E402
from oracle import render_oracle as oracle  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    s = b.abs().max().item()
    return (a - b).abs().max().item() / (s if s > 0 else 1.0), s


def main(names):
    for name in names:
        case = _cases.CASE_BY_NAME[name]
        run = _harness.oracle_run(case)
        st = run["out"]["stages"]
        gen = _cases.build_mirror(case, DEV)
        c = case.cfg
        B, n, s = case.batch, c["img_size"] ** 2, c["num_steps"]
        Cc = st["raw_coarse"].shape[-1]
        print("==== %s  B=%d n=%d s=%d C=%d" % (name, B, n, s, Cc))
        # ---------- composite backward vs fp64 autograd of the oracle's compositing ----------
        raw_c = st["raw_coarse"].double().requires_grad_(True)
        raw_f = st["raw_fine"].double().requires_grad_(True)
        all_raw = torch.cat([raw_f, raw_c], dim=-2)
        all_z = torch.cat([st["z_fine"], st["z_coarse"]], dim=-2).double()
        _, order = torch.sort(all_z, dim=-2)
        all_z = torch.gather(all_z, -2, order)
        all_raw = torch.gather(all_raw, -2, order.expand(-1, -1, -1, Cc))

        class D:
            log = []
            def randn(self, *shape):
                return run["draws"][-1][1].double()
        torch.set_default_dtype(torch.float64)
        px, _, _, _ = oracle.alpha_composite(all_raw, all_z, D(), c["nerf_noise"], c["clamp_mode"], last_back=c.get("last_back", False),
                                             white_back=c.get("white_back", False), black_back=c.get("black_back", False))
        torch.set_default_dtype(torch.float32)
        if gen.softmax_label:
            px = torch.cat([torch.softmax(px[..., :-3], -1), px[..., -3:]], -1)
        r = c["img_size"]
        px = px.reshape(B, r, r, -1).permute(0, 3, 1, 2) * 2 - 1
        W = _cases.loss_weights(px.shape).double()
        (px * W).sum().backward()
        rd = ops.make_render_desc(batch=B, img_size=r, num_steps=s, hierarchical=True, clamp_mode=c["clamp_mode"],
                                  nerf_noise=c["nerf_noise"], fov=c["fov"], last_back=c.get("last_back", False),
                                  white_back=c.get("white_back", False), black_back=c.get("black_back", False),
                                  softmax_label=gen.softmax_label)
        rc, rf = st["raw_coarse"].to(DEV).contiguous(), st["raw_fine"].to(DEV).contiguous()
        zc, zf = st["z_coarse"].to(DEV).contiguous(), st["z_fine"].to(DEV).contiguous()
        d_c, d_f = torch.empty_like(rc), torch.empty_like(rf)
        noise = run["draws"][-1][1].to(DEV).contiguous()
        _lib.check(_lib.lib().fenerf_composite_backward(C.byref(rd), Cc, rc.data_ptr(), zc.data_ptr(), rf.data_ptr(), zf.data_ptr(),
                                                        noise.data_ptr() if c["nerf_noise"] else 0, W.float().to(DEV).contiguous().data_ptr(),
                                                        d_c.data_ptr(), d_f.data_ptr(), 0))
        torch.cuda.synchronize()
        print("composite_backward  d_raw_c rel %.2e (scale %.3g)   d_raw_f rel %.2e (scale %.3g)" % (
            *rel(d_c.cpu().double(), raw_c.grad), *rel(d_f.cpu().double(), raw_f.grad)))
        # ---------- field backward vs fp32 autograd of the oracle's field on the GPU ----------
        film = run["film"].to(DEV).clone().requires_grad_(True)
        pts = st["points_fine"].reshape(B, n * s, 3).to(DEV)
        dirs = st["dirs"].to(DEV)
        dirs_pp = dirs.unsqueeze(-2).expand(-1, -1, s, -1).reshape(B, n * s, 3)
        torch.backends.cuda.matmul.allow_tf32 = False
        out = oracle.field_eval(gen.siren, pts, film, dirs_pp)
        R = torch.randn_like(out) * 0.01
        (out * R).sum().backward()
        want = {k: p.grad.clone() for k, p in gen.named_parameters() if p.grad is not None}
        want_film = film.grad.clone()
        with torch.no_grad():
            scale = torch.exp2(4.0 - torch.ceil(torch.log2(R.abs().max()))).float().reshape(1)
            fb = backward._FieldBackward(gen.siren, film.detach(), scale, (1.0 / scale).reshape(1))
            fb.add_points(pts.contiguous(), dirs.contiguous(), s, False, out.detach().contiguous(), R.contiguous())
            d_film, grads = fb.finish()
        print("film grad: freq rel %.2e (scale %.3g)  phase rel %.2e (scale %.3g)" % (
            *rel(d_film[:, :, 0], want_film[:, :, 0]), *rel(d_film[:, :, 1], want_film[:, :, 1])))
        for li in range(d_film.shape[1]):
            print("   layer %2d  freq rel %.2e  phase rel %.2e" % (li, rel(d_film[:, li, 0], want_film[:, li, 0])[0],
                                                                 rel(d_film[:, li, 1], want_film[:, li, 1])[0]))
        for k, p in gen.named_parameters():
            if id(p) in grads and k in ("siren." + kk for kk in []):
                pass
        for k, p in gen.siren.named_parameters():
            if id(p) in grads:
                g = grads[id(p)].reshape(p.shape)
                w = want["siren." + k]
                if w.numel() > 1e7:
                    print("   %-40s rel %.2e (scale %.3g)  [abs-sum %.4g vs %.4g]" % (k, *rel(g, w), g.abs().sum().item(), w.abs().sum().item()))
                else:
                    print("   %-40s rel %.2e (scale %.3g)" % (k, *rel(g, w)))


if __name__ == "__main__":
    main(sys.argv[1:] or ["a_small", "b_small"])
