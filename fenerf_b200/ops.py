"""Thin Python operators over the C-ABI (include/fenerf_b200.h).

Everything here is plumbing: validate tensors (device / dtype / contiguity, in the style of the
reference's own extension shim siren/op/fused_bias_act.cpp:7-9), hand raw device pointers and the
current CUDA stream to the library, wrap the outputs.  No arithmetic of the hot path happens in
Python or in torch ops.
"""
import ctypes as C
import math
import os

import torch

from . import _lib

_DEFAULT_PRECISION = os.environ.get("FENERF_B200_PRECISION", "guard")


def default_precision():
    return _DEFAULT_PRECISION


def set_default_precision(name):
    global _DEFAULT_PRECISION
    if name not in _lib.PRECISION:
        raise ValueError("precision must be one of %s" % sorted(_lib.PRECISION))
    _DEFAULT_PRECISION = name


def _precision_code(precision):
    name = precision if precision is not None else _DEFAULT_PRECISION
    if isinstance(name, int):
        return name
    return _lib.PRECISION[name]


def _chk(t, name, device=None, dtype=torch.float32):
    if t is None:
        return 0
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (fenerf_b200 has no CPU path)" % name)
    if device is not None and t.device != device:
        raise RuntimeError("%s is on %s, expected %s" % (name, t.device, device))
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    return t.data_ptr()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _prep(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def make_render_desc(*, batch, img_size, num_steps, hierarchical, clamp_mode, nerf_noise, fov, last_back=False,
                     white_back=False, black_back=False, fill_mode=None, fill_color="black", softmax_label=False,
                     lock_view_dependence=False, precision=None, guard_tau=0.0):
    if fill_mode not in _lib.FILL_MODE:
        raise ValueError("unknown fill_mode %r" % (fill_mode,))
    # the reference evaluates np.tan((2*pi*fov/360)/2) in double and divides a float32 tensor by it
    tan_half = math.tan((2 * math.pi * fov / 360) / 2)
    return _lib.RenderDesc(
        batch=batch, img_h=img_size, img_w=img_size, num_steps=num_steps, hierarchical=int(bool(hierarchical)),
        clamp_mode=_lib.CLAMP.get(clamp_mode, -1), last_back=int(bool(last_back)), white_back=int(bool(white_back)),
        black_back=int(bool(black_back)), fill_mode=_lib.FILL_MODE[fill_mode],
        fill_color=_lib.FILL_COLOR.get(fill_color, -1.0), softmax_label=int(bool(softmax_label)),
        lock_view_dependence=int(bool(lock_view_dependence)), precision=_precision_code(precision),
        noise_std=float(nerf_noise), tan_half_fov=tan_half, guard_tau=float(guard_tau))


# --------------------------------------------------------------------------------------------
# point network
# --------------------------------------------------------------------------------------------
POINTS_SIGMA_ONLY = 0x100     # FENERF_POINTS_SIGMA_ONLY


def siren_sigma(module, points, film, precision=None):
    """Density only: (B,P,3) points, (B,L,2,256) FiLM table -> (B,P,1).

    What extract_double_semantic_shapes.py:59-62 keeps of the point network's output
    (`coarse_output[:, :, -1:]` on a 256^3 grid); the colour and label branches are not evaluated.
    """
    b, p, _ = points.shape
    dirs = torch.zeros((b, 1, 3), dtype=torch.float32, device=points.device)
    out = siren_points(module, points, film, dirs, precision=precision, dir_group=p, _sigma_only=True)
    return out[..., -1:]


def siren_points(module, points, film, ray_directions, precision=None, dir_group=None, only_idx=None, _sigma_only=False):
    """(B,P,3) points, (B,L,2,256) FiLM table, (B,P,3) or (B,P/g,3) directions -> (B,P,C).

    The entry behind <SIREN>.forward_with_frequencies_phase_shifts (siren/siren.py:164-178,
    1509-1530).  Forward-only: differentiating through it is section 8f-1 of SURVEY.md.
    """
    if needs_grad(module, points, film):
        raise NotImplementedError(GRAD_MESSAGE)
    packed = module.packed()
    device = packed.device
    pts = _prep(points, device)
    flm = _prep(film, device)
    dirs = _prep(ray_directions, device)
    b, p, _ = pts.shape
    if dir_group is None:
        if dirs.shape[1] == p:
            dir_group = 1
        else:
            if p % dirs.shape[1]:
                raise ValueError("ray_directions (%d) does not divide the point count (%d)" % (dirs.shape[1], p))
            dir_group = p // dirs.shape[1]
    if flm.shape[0] != b or flm.shape[2:] != (2, _lib.HIDDEN):
        raise ValueError("film table has shape %s" % (tuple(flm.shape),))
    out = torch.empty((b, p, packed.desc.out_dim), dtype=torch.float32, device=device) if only_idx is None else only_idx[1]
    idx_ptr, n_only = 0, 0
    if only_idx is not None:
        idx = only_idx[0]
        idx_ptr, n_only = _chk(idx, "only_idx", device, torch.int32), idx.numel()
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_siren_points(
            C.byref(packed.desc), packed.ptr, _chk(pts, "points"), _chk(dirs, "ray_directions"), _chk(flm, "film"),
            b, p, dir_group, _precision_code(precision) | (POINTS_SIGMA_ONLY if _sigma_only else 0), idx_ptr, n_only,
            _chk(out, "out"), _stream(device)))
    return out


GRAD_MESSAGE = ("fenerf_b200: the point-network entry (<SIREN>.forward / forward_with_frequencies_phase_shifts) is "
                "forward-only; differentiate through the generator's forward / forward_with_frequencies "
                "(fenerf_b200/backward.py), or wrap the call in torch.no_grad()")


def needs_grad(module, *tensors):
    """True when the caller expects autograd to flow through the point network."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


# --------------------------------------------------------------------------------------------
# render stages (exposed for stage-level parity tests and for callers that bring their own rays)
# --------------------------------------------------------------------------------------------
def camera_poses(n, mode, h_stddev, v_stddev, h_mean, v_mean, rng, device):
    """Camera pose sampling + look-at matrix in one launch (sample_camera_positions + create_cam2world_matrix,
    generators/volumetric_rendering.py:170-248).  The draws are made here with `rng` in the reference's order
    (theta then phi; 'hybrid' first flips Python's `random.random()`); every mode the reference names is
    covered, anything else means "the mean pose" as in the reference's else-branch.
    Returns (cam2world (n,4,4), pitch (n,1), yaw (n,1))."""
    code = _lib.CAMERA_MODE.get(mode, 0)
    h_stddev, v_stddev, h_mean, v_mean = float(h_stddev), float(v_stddev), float(h_mean), float(v_mean)
    if mode == "hybrid":
        if rng.coin() < 0.5:
            code, h_stddev, v_stddev = 1, 2 * h_stddev, 2 * v_stddev
        else:
            code = 2
    d_theta = d_phi = None
    if code in (1, 4):
        d_theta, d_phi = rng.rand(n, 1), rng.rand(n, 1)
    elif code == 2:
        d_theta, d_phi = rng.randn(n, 1), rng.randn(n, 1)
    elif code == 3:
        d_theta, d_phi = rng.randn(n, 1, 4), rng.randn(n, 1, 4)
    if code == 4:
        v_stddev, v_mean = v_stddev / math.pi, v_mean / math.pi
    c2w = torch.empty((n, 4, 4), dtype=torch.float32, device=device)
    pitch = torch.empty((n, 1), dtype=torch.float32, device=device)
    yaw = torch.empty((n, 1), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_camera_poses(
            n, code, h_stddev, v_stddev, h_mean, v_mean,
            _chk(d_theta.contiguous(), "draw_theta", device) if d_theta is not None else 0,
            _chk(d_phi.contiguous(), "draw_phi", device) if d_phi is not None else 0,
            c2w.data_ptr(), pitch.data_ptr(), yaw.data_ptr(), _stream(device)))
    return c2w, pitch, yaw


_TABLES = {}


def ray_tables(img_size, num_steps, ray_start, ray_end, device):
    """Cached linspace tables (torch.linspace, so the values are the reference's to the bit)."""
    key = (img_size, num_steps, float(ray_start), float(ray_end), str(device))
    t = _TABLES.get(key)
    if t is None:
        from .generators import volumetric_rendering as vr
        t = vr.ray_tables(img_size, num_steps, ray_start, ray_end, device)
        if len(_TABLES) > 64:
            _TABLES.clear()
        _TABLES[key] = t
    return t


def ray_setup(rd, x_lin, y_lin, z_lin, cam2world, rng_perturb):
    device = cam2world.device
    b, n, s = rd.batch, rd.img_h * rd.img_w, rd.num_steps
    points = torch.empty((b, n, s, 3), dtype=torch.float32, device=device)
    z_vals = torch.empty((b, n, s, 1), dtype=torch.float32, device=device)
    dirs = torch.empty((b, n, 3), dtype=torch.float32, device=device)
    origins = torch.empty((b, 3), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_ray_setup(
            C.byref(rd), _chk(x_lin, "x_lin", device), _chk(y_lin, "y_lin", device), _chk(z_lin, "z_lin", device),
            _chk(cam2world, "cam2world", device), _chk(rng_perturb, "rng_perturb", device),
            points.data_ptr(), z_vals.data_ptr(), dirs.data_ptr(), origins.data_ptr(), _stream(device)))
    return points, z_vals, dirs, origins


def resample(rd, raw_coarse, z_vals, dirs, origins, rng_noise, rng_u, want_inds=False):
    device = raw_coarse.device
    b, n, s = rd.batch, rd.img_h * rd.img_w, rd.num_steps
    c = raw_coarse.shape[-1]
    z_fine = torch.empty((b, n, s, 1), dtype=torch.float32, device=device)
    pts = torch.empty((b, n, s, 3), dtype=torch.float32, device=device)
    inds = torch.empty((b * n, s), dtype=torch.int64, device=device) if want_inds else None
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_resample(
            C.byref(rd), c, _chk(raw_coarse, "raw_coarse", device), _chk(z_vals, "z_vals", device),
            _chk(dirs, "dirs", device), _chk(origins, "origins", device), _chk(rng_noise, "rng_noise", device),
            _chk(rng_u, "rng_u", device), z_fine.data_ptr(), pts.data_ptr(),
            inds.data_ptr() if inds is not None else 0, _stream(device)))
    return z_fine, pts, inds


def composite(rd, raw_coarse, z_coarse, raw_fine=None, z_fine=None, rng_noise=None, want_weights=False,
              want_sort_idx=False):
    device = raw_coarse.device
    b, n, s = rd.batch, rd.img_h * rd.img_w, rd.num_steps
    c = raw_coarse.shape[-1]
    ns = 2 * s if rd.hierarchical else s
    pad = rd.fill_mode in (_lib.FILL_MODE["seg_padding_background"], _lib.FILL_MODE["eval_seg_padding_background"])
    c_img = c - 1 + (1 if pad else 0)
    pixels = torch.empty((b, c_img, rd.img_h, rd.img_w), dtype=torch.float32, device=device)
    depth = torch.empty((b, n, 1), dtype=torch.float32, device=device)
    wsum = torch.empty((b, n, 1), dtype=torch.float32, device=device)
    weights = torch.empty((b, n, ns, 1), dtype=torch.float32, device=device) if want_weights else None
    sidx = torch.empty((b, n, ns), dtype=torch.int32, device=device) if want_sort_idx else None
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_composite(
            C.byref(rd), c, _chk(raw_coarse, "raw_coarse", device), _chk(z_coarse, "z_coarse", device),
            _chk(raw_fine, "raw_fine", device), _chk(z_fine, "z_fine", device), _chk(rng_noise, "rng_noise", device),
            pixels.data_ptr(), depth.data_ptr(), wsum.data_ptr(), weights.data_ptr() if weights is not None else 0,
            sidx.data_ptr() if sidx is not None else 0, _stream(device)))
    return pixels, depth, wsum, weights, sidx


_WORKSPACES = {}


def _workspace(device, nbytes):
    """One grow-only scratch buffer per (device, stream): the C-ABI never allocates."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.05) + 256, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


DEFAULT_GUARD_TAU = 1.5e-3


def guard_stats(device):
    """The GUARD self-check of the last render_forward on the current stream (fenerf_guard_stats): how far the tcgen05
    far-sample densities were from their fp32 re-evaluation.  Synchronises the stream."""
    device = torch.device(device)
    ws = _WORKSPACES.get((device, torch.cuda.current_stream(device).cuda_stream))
    if ws is None:
        return None
    rep = _lib.GuardReport()
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_guard_stats((ws.data_ptr() + 255) // 256 * 256, C.byref(rep), _stream(device)))
    return dict(refined=rep.refined, max_abs_delta=rep.max_abs_delta, sign_flips=rep.sign_flips, tau=rep.tau)


def render_forward(module, rd, film, x_lin, y_lin, z_lin, cam2world, rng_perturb, rng_noise_c, rng_u, rng_noise_f,
                   want_depth=True, want_weights_sum=True, want_weights=False, want_inds=False):
    """One call into fenerf_render_forward: the whole render after the mapping network."""
    packed = module.packed()
    device = packed.device
    lib = _lib.lib()
    b, n, s = rd.batch, rd.img_h * rd.img_w, rd.num_steps
    ns = 2 * s if rd.hierarchical else s
    c = packed.desc.out_dim
    pad = rd.fill_mode in (_lib.FILL_MODE["seg_padding_background"], _lib.FILL_MODE["eval_seg_padding_background"])
    c_img = c - 1 + (1 if pad else 0)
    film = _prep(film, device)
    if film.shape != (b, packed.desc.trunk_layers + packed.desc.color_layers, 2, _lib.HIDDEN):
        raise ValueError("film table has shape %s" % (tuple(film.shape),))
    pixels = torch.empty((b, c_img, rd.img_h, rd.img_w), dtype=torch.float32, device=device)
    depth = torch.empty((b, n, 1), dtype=torch.float32, device=device) if want_depth else None
    wsum = torch.empty((b, n, 1), dtype=torch.float32, device=device) if want_weights_sum else None
    weights = torch.empty((b, n, ns, 1), dtype=torch.float32, device=device) if want_weights else None
    inds = torch.empty((b * n, s), dtype=torch.int64, device=device) if (want_inds and rd.hierarchical) else None
    with torch.cuda.device(device):
        nbytes = lib.fenerf_workspace_bytes(C.byref(rd), C.byref(packed.desc))
        ws = _workspace(device, nbytes)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        _lib.check(lib.fenerf_render_forward(
            C.byref(rd), C.byref(packed.desc), packed.ptr, _chk(film, "film", device),
            _chk(x_lin, "x_lin", device), _chk(y_lin, "y_lin", device), _chk(z_lin, "z_lin", device),
            _chk(cam2world, "cam2world", device), _chk(rng_perturb, "rng_perturb", device),
            _chk(rng_noise_c, "rng_noise_c", device), _chk(rng_u, "rng_u", device),
            _chk(rng_noise_f, "rng_noise_f", device),
            pixels.data_ptr(), depth.data_ptr() if depth is not None else 0,
            wsum.data_ptr() if wsum is not None else 0, weights.data_ptr() if weights is not None else 0,
            inds.data_ptr() if inds is not None else 0, ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()),
            _stream(device)))
    return pixels, depth, wsum, weights, inds


def render_forward_stages(module, rd, film, x_lin, y_lin, z_lin, cam2world, rng_perturb, rng_noise_c, rng_u, rng_noise_f):
    """fenerf_render_forward into a PRIVATE workspace, returned together with typed views of the intermediates
    it leaves there (fenerf_workspace_layout): what the backward consumes (fenerf_b200/backward.py)."""
    packed = module.packed()
    device = packed.device
    lib = _lib.lib()
    b, n, s = rd.batch, rd.img_h * rd.img_w, rd.num_steps
    c = packed.desc.out_dim
    film = _prep(film, device)
    pixels = torch.empty((b, c - 1, rd.img_h, rd.img_w), dtype=torch.float32, device=device)
    off = _lib.WorkspaceOffsets()
    with torch.cuda.device(device):
        _lib.check(lib.fenerf_workspace_layout(C.byref(rd), C.byref(packed.desc), C.byref(off)))
        ws = torch.empty(off.total + 256, dtype=torch.uint8, device=device)
        base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
        _lib.check(lib.fenerf_render_forward(
            C.byref(rd), C.byref(packed.desc), packed.ptr, _chk(film, "film", device),
            _chk(x_lin, "x_lin", device), _chk(y_lin, "y_lin", device), _chk(z_lin, "z_lin", device),
            _chk(cam2world, "cam2world", device), _chk(rng_perturb, "rng_perturb", device),
            _chk(rng_noise_c, "rng_noise_c", device), _chk(rng_u, "rng_u", device), _chk(rng_noise_f, "rng_noise_f", device),
            pixels.data_ptr(), 0, 0, 0, 0, ws.data_ptr() + base, ws.numel() - base, _stream(device)))

    def view(offset, shape):
        numel = 1
        for d in shape:
            numel *= d
        return ws[base + offset: base + offset + numel * 4].view(torch.float32).view(shape)

    st = dict(pixels=pixels, workspace=ws, points_c=view(off.points_coarse, (b, n, s, 3)), z_c=view(off.z_coarse, (b, n, s)),
              dirs=view(off.dirs, (b, n, 3)), raw_c=view(off.raw_coarse, (b, n, s, c)), raw_f=None, z_f=None, points_f=None)
    if rd.hierarchical:
        st.update(points_f=view(off.points_fine, (b, n, s, 3)), z_f=view(off.z_fine, (b, n, s)),
                  raw_f=view(off.raw_fine, (b, n, s, c)))
    return st


def mapping_film(net, z, film, first_layer, n_layers, avg=None, psi=1.0):
    """fenerf_mapping_film: CustomMappingNetwork + `15 f + 30` (+ psi truncation towards `avg` = (avg_frequencies,
    avg_phase_shifts)) written straight into layers [first_layer, first_layer + n_layers) of the FiLM table
    `film` (B, L, 2, 256).  Two launches instead of ~15 (no_grad callers only)."""
    linears = [m for m in net.network if isinstance(m, torch.nn.Linear)]
    if len(linears) != 5:
        raise ValueError("mapping network: expected 5 Linear layers, got %d" % len(linears))
    device = film.device
    p = _lib.MappingParams()
    keep = []
    for i, lin in enumerate(linears):
        w, b = _prep(lin.weight, device), _prep(lin.bias, device)
        keep += [w, b]
        p.weight[i], p.bias[i] = w.data_ptr(), b.data_ptr()
    p.z_dim, p.hidden_dim = linears[0].in_features, linears[0].out_features
    z = _prep(z, device)
    bsz = z.shape[0]
    h = torch.empty((min(bsz, 32), 256), dtype=torch.float32, device=device)
    af = ap = None
    if avg is not None:
        af, ap = _prep(avg[0].reshape(-1), device), _prep(avg[1].reshape(-1), device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().fenerf_mapping_film(
            C.byref(p), _chk(z, "z", device), bsz, n_layers, first_layer, film.shape[1], _chk(af, "avg_frequencies", device),
            _chk(ap, "avg_phase_shifts", device), float(psi), h.data_ptr(), _chk(film, "film", device), _stream(device)))
    return film


# --------------------------------------------------------------------------------------------
# tcgen05 GEMMs of the backward (csrc/gemm5.cu)
# --------------------------------------------------------------------------------------------
def gemm_nt(a16, b16, out_dtype=torch.float32, gate=None):
    """(M, 256) fp16 . (256, 256)^T fp16 -> (M, 256) fp32 or fp16 (fenerf_gemm_nt_f16); `gate` (M, 256) fp16 multiplies
    the fp16 output in the epilogue."""
    dev = a16.device
    m = a16.shape[0]
    out = torch.empty((m, 256), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fenerf_gemm_nt_f16(
            _chk(a16, "A", dev, torch.float16), _chk(b16, "B", dev, torch.float16), m,
            out.data_ptr() if out_dtype == torch.float32 else 0, out.data_ptr() if out_dtype == torch.float16 else 0,
            _chk(gate, "gate", dev, torch.float16), _stream(dev)))
    return out


def gemm_nt_film(a16, w16, bias, film, b0, layer, ppb, narrow_in=None, narrow_w=None):
    """One FiLM layer's recompute with the epilogue fused (fenerf_gemm_nt_film): -> (a, gate), both (M, 256) fp16.
    narrow_in (M, 64) / narrow_w (256, 64) fp16 (zero padded): extra inputs of the layer."""
    dev = a16.device
    m = a16.shape[0]
    a_out = torch.empty((m, 256), dtype=torch.float16, device=dev)
    g_out = torch.empty((m, 256), dtype=torch.float16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fenerf_gemm_nt_film(
            _chk(a16, "A", dev, torch.float16), _chk(w16, "W", dev, torch.float16), m, _chk(bias, "bias", dev),
            film[b0, layer].data_ptr(), film.stride(0), ppb, _chk(narrow_in, "narrow_in", dev, torch.float16),
            _chk(narrow_w, "narrow_w", dev, torch.float16), a_out.data_ptr(), g_out.data_ptr(), _stream(dev)))
    return a_out, g_out


def gemm_tn(x16, y16, batch, ppb, slices=None, colsum=False):
    """Per image b: X_b^T Y_b with X, Y (batch * ppb, 256) fp16 -> (batch, 256, 256) fp32 (fenerf_gemm_tn_f16; the
    split-K partials of the CTAs are summed here).  colsum=True also returns the column sums of X per image (batch, 256)."""
    dev = x16.device
    if slices is None:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        slices = max(1, min((ppb + 63) // 64, (sms + batch - 1) // batch))
    partial = torch.empty((batch, slices, 256, 256), dtype=torch.float32, device=dev)
    cs = torch.empty((batch, slices, 256), dtype=torch.float32, device=dev) if colsum else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fenerf_gemm_tn_f16(_chk(x16, "X", dev, torch.float16), _chk(y16, "Y", dev, torch.float16), batch,
                                                 ppb, slices, partial.data_ptr(), cs.data_ptr() if colsum else 0, _stream(dev)))
    out = partial.sum(1) if slices > 1 else partial[:, 0]
    return (out, cs.sum(1)) if colsum else out
