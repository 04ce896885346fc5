"""Backward of the render: a ``torch.autograd.Function`` over the CUDA library.

What the reference differentiates (train_double_latent_semantic.py:405-446, the G step under autocast;
inverse_render_double_semantic.py:385-407, Adam on FiLM offsets through ``forward_with_frequencies``,
generators/generators.py:735-798): the final ``fancy_integration`` over the merged samples and both
point-network passes.  Ray set-up and resampling are ``no_grad`` there too (generators.py:41, 59).

forward   ONE ``fenerf_render_forward`` call into a private workspace -- the same kernels, numerics and
          precision modes as the no_grad path; the workspace (sample points, depths, raw field outputs of
          both passes) is what the backward needs and is kept alive by the autograd node.
backward  ``fenerf_composite_backward`` (d pixels -> d raw outputs, one warp per ray), then the point
          network layer by layer, recomputing activations chunk by chunk (nothing but the workspace
          survives from the forward):
            recompute   ``fenerf_gemm_nt_film``: z = a W^T on tcgen05 with the epilogue fused -- a = sin(f z + p) and
                        the gate f cos(f z + p) leave as fp16, z never does
            backward    ``fenerf_gate_backward`` dZ = dA * gate (+ per-image column sums), dA' = dZ W
                        (``fenerf_gemm_nt_f16``) and the per-image dW_b = dZ^T a (``fenerf_gemm_tn_f16``, split-K)
          FiLM gradients need no further pass over the points:  dp = db_b / f,
          df = (sum_k W[f,k] dW_b[f,k]) / f + b dp   (u = f z + p, z = W a + b).
          Heads, the pre-multiplied label chain and the grid (``fenerf_grid_scatter_add``) close the chain.
Gradients flow to the FiLM table (and through torch's autograd into the mapping network / latents /
frequency offsets) and to every field parameter.  The fp16 gradient stream is scaled by a power of two
taken from max|d raw| on the device (no host sync) and unscaled at the end.

The three 256-wide products per layer run on tcgen05 (csrc/gemm5.cu: ``fenerf_gemm_nt_film`` -- the recompute with its FiLM
epilogue fused, ``fenerf_gemm_nt_f16`` for dA' = dZ W, ``fenerf_gemm_tn_f16`` split-K for the per-image dW); only the narrow
products (heads, the 3 / 35-wide inputs) go to the library.  ``FENERF_B200_BWD_GEMM=cublas`` switches the wide ones back
(A/B timing); ``precision='exact'`` always uses fp32 library GEMMs.  (The kernels can also fold the next layer's gate multiply
into the dA product's epilogue and produce the bias column sums from the dW kernel's staged tiles -- measured: the gate kernel's
17 ms disappear but the two GEMMs slow down by as much, both being HBM-bound; the chain below keeps the separate gate kernel.)
"""
import ctypes as C

import torch

from . import _lib, ops, packing

import os

CHUNK_POINTS = 1 << 19
#: the 256-wide products: 'tcgen05' = csrc/gemm5.cu (default), 'cublas' = torch.mm / bmm (kept for A/B timing and as the
#: fp32 path of precision='exact')
BWD_GEMM = os.environ.get("FENERF_B200_BWD_GEMM", "tcgen05")


def _ptr(t):
    return t.data_ptr() if t is not None else 0


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _mm32(a, b):
    """fp16 x fp16 -> fp32 (tensor cores, fp32 accumulate AND fp32 output); plain fp32 in the exact mode."""
    if a.dtype == torch.float32:
        return torch.mm(a, b)
    return torch.mm(a, b, out_dtype=torch.float32)


def _bmm32(a, b):
    if a.dtype == torch.float32:
        return torch.bmm(a, b)
    return torch.bmm(a, b, out_dtype=torch.float32)


class FieldWeights:
    """The field's parameters in the roles the backward needs, plus fp16 copies for the GEMMs."""

    def __init__(self, module):
        spec = module.field_spec()
        self.spec = spec
        net = list(module.network)
        color = module.color_layer_sine
        color = list(color) if isinstance(color, torch.nn.ModuleList) else [color]
        self.trunk = [(l.layer.weight, l.layer.bias) for l in net]
        self.color = [(l.layer.weight, l.layer.bias) for l in color]
        self.sigma = (module.final_layer.weight, module.final_layer.bias)
        self.rgb = (module.color_layer_linear[0].weight, module.color_layer_linear[0].bias)
        self.labels = []
        if spec.label_dim:
            self.labels = [(m.weight, m.bias) for m in module.label_layer_linear if isinstance(m, torch.nn.Linear)]
        self.grid = module.spatial_embeddings if spec.grid_channels else None

    def parameters(self):
        ps = []
        for w, b in self.trunk + self.color + [self.sigma, self.rgb] + self.labels:
            ps += [w, b]
        if self.grid is not None:
            ps.append(self.grid)
        return ps


def _label_eff(labels):
    """Weff, beff of the activation-free label chain (siren.py:1486-1490 / 1189-1191), fp32 via fp64."""
    ws = [w.detach().double() for w, _ in labels]
    bs = [b.detach().double() for _, b in labels]
    if len(labels) == 3:
        u = ws[2] @ ws[1]
        return (u @ ws[0]).float(), (u @ bs[0] + ws[2] @ bs[1] + bs[2]).float()
    return (ws[1] @ ws[0]).float(), (ws[1] @ bs[0] + bs[1]).float()


def _label_chain_grads(labels, d_weff, d_beff):
    """Gradients of the chain's own parameters from those of the pre-multiplied map."""
    ws = [w.detach().double() for w, _ in labels]
    bs = [b.detach().double() for _, b in labels]
    dw, db = d_weff.double(), d_beff.double()
    if len(labels) == 3:
        w1, w2, w3 = ws
        b1, b2, _ = bs
        u = w3 @ w2
        g1w, g1b = u.t() @ dw, u.t() @ db
        g2w = w3.t() @ (dw @ w1.t() + torch.outer(db, b1))
        g2b = w3.t() @ db
        g3w = dw @ (w2 @ w1).t() + torch.outer(db, w2 @ b1 + b2)
        return [(g1w, g1b), (g2w, g2b), (g3w, db)]
    w1, w3 = ws
    b1, _ = bs
    return [(w3.t() @ dw, w3.t() @ db), (dw @ w1.t() + torch.outer(db, b1), db)]


class _FieldBackward:
    """Accumulates the gradients of one field over any number of point sets."""

    def __init__(self, module, film, scale, inv_scale, exact=False):
        self.module = module
        # stream element type: fp16 (default) or fp32 (parity mode, with precision='exact': plain fp32 GEMMs)
        self.dt = torch.float32 if exact else torch.float16
        self.dtc = 1 if exact else 0
        self.fw = FieldWeights(module)
        self.spec = self.fw.spec
        self.packed = module.packed()
        self.film = film.detach().float().contiguous()           # (B, n_film, 2, 256)
        self.scale, self.inv_scale = scale, inv_scale
        dev = self.film.device
        self.dev = dev
        B = self.film.shape[0]
        T, Cn = len(self.fw.trunk), len(self.fw.color)
        self.T, self.Cn, self.n_film = T, Cn, T + Cn
        G = self.spec.grid_channels
        self.kx = 3 + G
        self.kx_pad = (self.kx + 7) // 8 * 8
        # per-image accumulators (scaled by `scale`)
        self.colsum = torch.zeros((B, self.n_film, 256), dtype=torch.float32, device=dev)
        self.dW_b = [None] * self.n_film
        self.dW_b[0] = torch.zeros((B, 256, 8), dtype=torch.float32, device=dev)
        for i in range(1, self.n_film):
            self.dW_b[i] = torch.zeros((B, 256, 256), dtype=torch.float32, device=dev)
        self.dWx_b = torch.zeros((B, 256, self.kx_pad), dtype=torch.float32, device=dev)   # colour 0's narrow inputs
        self.d_heads_w = torch.zeros((32, 256), dtype=torch.float32, device=dev)
        self.d_heads_b = torch.zeros((32,), dtype=torch.float32, device=dev)
        self.d_rgb_w = torch.zeros((8, 256), dtype=torch.float32, device=dev)
        self.d_rgb_b = torch.zeros((8,), dtype=torch.float32, device=dev)
        self.grid_grad_cl = None
        if G:
            r = self.spec.grid_res
            self.grid_grad_cl = torch.zeros((r, r, r, G), dtype=torch.float32, device=dev)
        # fp16 / fp32 weight views for the GEMMs
        fw = self.fw
        self.W0 = fw.trunk[0][0].detach().float().contiguous()                    # (256, 3)
        self.own_gemm = (not exact) and BWD_GEMM == "tcgen05"
        self.Wh16 = [None] + [w.detach().to(self.dt).contiguous() for w, _ in fw.trunk[1:]]
        wc0 = fw.color[0][0].detach().float()
        self.Wc0x_narrow = wc0[:, :self.kx].contiguous()                          # (256, 3 + G) fp32
        self.Wc16 = [wc0[:, self.kx:].to(self.dt).contiguous()] + [w.detach().to(self.dt).contiguous() for w, _ in fw.color[1:]]
        self.Wfeat16 = wc0[:, 3:self.kx].to(self.dt).contiguous() if G else None       # (256, G)
        # dA' = dZ W on the tcgen05 NT kernel wants the transposed weights as its (N, K) operand
        self.WhT16 = [None] + [w.t().contiguous() for w in self.Wh16[1:]] if self.own_gemm else None
        self.WcT16 = [w.t().contiguous() for w in self.Wc16] if self.own_gemm else None
        if self.own_gemm:
            self.Wn64 = torch.zeros((256, 64), dtype=torch.float16, device=dev)
            self.Wn64[:, :self.kx] = self.Wc0x_narrow
        L = self.spec.label_dim
        self.L = L
        heads = torch.zeros((32, 256), dtype=torch.float32, device=dev)
        if L:
            weff, _ = _label_eff(fw.labels)
            heads[:L] = weff
        heads[L] = fw.sigma[0].detach().float().reshape(-1)
        self.Wheads32 = heads
        rgbw = torch.zeros((8, 256), dtype=self.dt, device=dev)
        rgbw[:3] = fw.rgb[0].detach().to(self.dt)
        self.Wrgb16 = rgbw
        self.bias = [b.detach().float().contiguous() for _, b in fw.trunk + fw.color]

    # ---- one point set: points (B, ppb, 3), dirs (B, ppb/dir_group, 3), raw / d_raw (B, ppb, C) ----
    def add_points(self, points, dirs, dir_group, lock_dirs, raw, d_raw):
        B, ppb, _ = points.shape
        if ppb <= CHUNK_POINTS:
            k = max(1, CHUNK_POINTS // ppb)
            for b0 in range(0, B, k):
                b1 = min(B, b0 + k)
                self._chunk(points[b0:b1], dirs[b0:b1], dir_group, lock_dirs, raw[b0:b1], d_raw[b0:b1], b0, b1)
        else:
            step = CHUNK_POINTS // dir_group * dir_group
            for b in range(B):
                for p0 in range(0, ppb, step):
                    p1 = min(ppb, p0 + step)
                    self._chunk(points[b:b + 1, p0:p1], dirs[b:b + 1, p0 // dir_group:p1 // dir_group], dir_group, lock_dirs,
                                raw[b:b + 1, p0:p1], d_raw[b:b + 1, p0:p1], b, b + 1)

    def _stash(self, z, idx, b0, P, ppb, xin=None, wx=None):
        a = torch.empty((P, 256), dtype=self.dt, device=self.dev)
        g = torch.empty((P, 256), dtype=self.dt, device=self.dev)
        film_l = self.film[b0, idx]
        kx = 0 if xin is None else xin.shape[1]
        _lib.check(_lib.lib().fenerf_film_forward_stash(
            _ptr(z), self.bias[idx].data_ptr(), film_l.data_ptr(), self.film.stride(0), P, ppb,
            _ptr(xin), kx, _ptr(wx), a.data_ptr(), g.data_ptr(), self.dtc, _stream(self.dev)))
        return a, g

    def _gate(self, dA, gate, idx, b0, b1, P, ppb):
        cs = self.colsum[b0:b1, idx]
        tmp = torch.zeros((b1 - b0, 256), dtype=torch.float32, device=self.dev)
        _lib.check(_lib.lib().fenerf_gate_backward(dA.data_ptr(), gate.data_ptr(), P, ppb, tmp.data_ptr(), self.dtc, _stream(self.dev)))
        cs += tmp

    def _chunk(self, points, dirs, dir_group, lock_dirs, raw, d_raw, b0, b1):
        lib = _lib.lib()
        dev, spec = self.dev, self.spec
        k, ppb = points.shape[0], points.shape[1]
        P = k * ppb
        T, Cn = self.T, self.Cn
        points = points.contiguous()
        dirs = dirs.contiguous()
        raw = raw.contiguous()
        d_raw = d_raw.contiguous()
        with torch.cuda.device(dev):
            st = _stream(dev)
            # ---- recompute the forward, stashing activations and gates (fp16) ----
            x = (points.reshape(P, 3) * spec.input_scale).contiguous() if spec.input_scale != 1.0 else points.reshape(P, 3)
            extras = torch.empty((P, self.kx), dtype=torch.float32, device=dev)
            _lib.check(lib.fenerf_extras_gather(C.byref(self.packed.desc), self.packed.ptr, points.data_ptr(), dirs.data_ptr(),
                                                P, ppb, dir_group, int(bool(lock_dirs)), extras.data_ptr(), st))
            A, Gt = [None] * self.n_film, [None] * self.n_film
            own = self.own_gemm
            A[0], Gt[0] = self._stash(None, 0, b0, P, ppb, xin=x, wx=self.W0)
            for l in range(1, T):
                if own:     # z = a W^T with the FiLM epilogue fused: z never leaves the SM
                    A[l], Gt[l] = ops.gemm_nt_film(A[l - 1], self.Wh16[l], self.bias[l], self.film, b0, l, ppb)
                else:
                    A[l], Gt[l] = self._stash(_mm32(A[l - 1], self.Wh16[l].t()), l, b0, P, ppb)
            if own:     # the narrow inputs [dir, grid features] ride as a fifth 64-wide k-chunk of the same kernel
                e64 = torch.zeros((P, 64), dtype=torch.float16, device=dev)
                e64[:, :self.kx] = extras
                A[T], Gt[T] = ops.gemm_nt_film(A[T - 1], self.Wc16[0], self.bias[T], self.film, b0, T, ppb, narrow_in=e64,
                                               narrow_w=self.Wn64)
                del e64
            else:
                A[T], Gt[T] = self._stash(_mm32(A[T - 1], self.Wc16[0].t()), T, b0, P, ppb, xin=extras, wx=self.Wc0x_narrow)
            for j in range(1, Cn):
                if own:
                    A[T + j], Gt[T + j] = ops.gemm_nt_film(A[T + j - 1], self.Wc16[j], self.bias[T + j], self.film, b0, T + j, ppb)
                else:
                    A[T + j], Gt[T + j] = self._stash(_mm32(A[T + j - 1], self.Wc16[j].t()), T + j, b0, P, ppb)
            # ---- head gradients ----
            dH = torch.empty((P, 32), dtype=self.dt, device=dev)
            dRGB = torch.empty((P, 8), dtype=self.dt, device=dev)
            _lib.check(lib.fenerf_head_grads(d_raw.data_ptr(), raw.data_ptr(), P, spec.out_dim, self.L, self.scale.data_ptr(),
                                             dH.data_ptr(), dRGB.data_ptr(), self.dtc, st))
            a_last = A[self.n_film - 1]
            self.d_rgb_w += _mm32(dRGB.t(), a_last)
            self.d_rgb_b += dRGB.float().sum(0)
            self.d_heads_w += _mm32(dH.t(), A[T - 1])
            self.d_heads_b += dH.float().sum(0)
            # ---- colour branch, top down ----
            dA = torch.mm(dRGB, self.Wrgb16)                                   # (P, 256) fp16
            for j in range(Cn - 1, -1, -1):
                idx = T + j
                self._gate(dA, Gt[idx], idx, b0, b1, P, ppb)                   # dA is dZ now
                a_in = A[idx - 1]
                dz3 = dA.view(k, ppb, 256).transpose(1, 2)
                self.dW_b[idx][b0:b1] += ops.gemm_tn(dA, a_in, k, ppb) if own else _bmm32(dz3, a_in.view(k, ppb, 256))
                if j == 0:
                    e16 = torch.zeros((P, self.kx_pad), dtype=self.dt, device=dev)
                    e16[:, :self.kx] = extras
                    self.dWx_b[b0:b1] += _bmm32(dz3, e16.view(k, ppb, self.kx_pad))
                    if spec.grid_channels:
                        d_feat = torch.mm(dA, self.Wfeat16).contiguous()       # (P, G) fp16
                        _lib.check(lib.fenerf_grid_scatter_add(C.byref(self.packed.desc), points.data_ptr(), d_feat.data_ptr(),
                                                               d_feat.shape[1], P, self.grid_grad_cl.data_ptr(), self.dtc, st))
                dA = ops.gemm_nt(dA, self.WcT16[j], torch.float16) if own else torch.mm(dA, self.Wc16[j])
                A[idx], Gt[idx] = None, None
            # ---- trunk: colour-branch gradient + sigma / label heads ----
            dA += torch.mm(dH.float(), self.Wheads32)
            for l in range(T - 1, 0, -1):
                self._gate(dA, Gt[l], l, b0, b1, P, ppb)
                if own:
                    self.dW_b[l][b0:b1] += ops.gemm_tn(dA, A[l - 1], k, ppb)
                    dA = ops.gemm_nt(dA, self.WhT16[l], torch.float16)
                else:
                    self.dW_b[l][b0:b1] += _bmm32(dA.view(k, ppb, 256).transpose(1, 2), A[l - 1].view(k, ppb, 256))
                    dA = torch.mm(dA, self.Wh16[l])
                A[l], Gt[l] = None, None
            self._gate(dA, Gt[0], 0, b0, b1, P, ppb)
            x16 = torch.zeros((P, 8), dtype=self.dt, device=dev)
            x16[:, :3] = x
            self.dW_b[0][b0:b1] += _bmm32(dA.view(k, ppb, 256).transpose(1, 2), x16.view(k, ppb, 8))

    # ---- after every point set: fold the per-image accumulators into parameter / FiLM gradients ----
    def finish(self):
        fw, inv = self.fw, self.inv_scale
        film = self.film
        d_film = torch.zeros_like(film)
        grads = {}
        layers = fw.trunk + fw.color
        for idx, (w, b) in enumerate(layers):
            f = film[:, idx, 0]                                             # (B, 256)
            db_b = self.colsum[:, idx]
            dp = db_b / f
            w32 = w.detach().float()
            if idx == 0:
                dwb = self.dW_b[0][:, :, :3]
                full = dwb
            elif idx == self.T:
                full = torch.cat([self.dWx_b[:, :, :self.kx], self.dW_b[idx]], dim=2)   # column order of the reference: [dir, feat, x]
            else:
                full = self.dW_b[idx]
            df = torch.einsum('fk,bfk->bf', w32, full) / f + b.detach().float().unsqueeze(0) * dp
            d_film[:, idx, 0] = df * inv
            d_film[:, idx, 1] = dp * inv
            grads[id(w)] = full.sum(0) * inv
            grads[id(b)] = db_b.sum(0) * inv
        L = self.L
        grads[id(fw.sigma[0])] = (self.d_heads_w[L] * inv).reshape(fw.sigma[0].shape)
        grads[id(fw.sigma[1])] = (self.d_heads_b[L] * inv).reshape(fw.sigma[1].shape)
        grads[id(fw.rgb[0])] = self.d_rgb_w[:3] * inv
        grads[id(fw.rgb[1])] = self.d_rgb_b[:3] * inv
        if L:
            chain = _label_chain_grads(fw.labels, self.d_heads_w[:L] * inv, self.d_heads_b[:L] * inv)
            for (w, b), (gw, gb) in zip(fw.labels, chain):
                grads[id(w)] = gw.float()
                grads[id(b)] = gb.float()
        if fw.grid is not None:
            out = torch.empty_like(fw.grid, dtype=torch.float32)
            _lib.check(_lib.lib().fenerf_grid_unpack_grad(C.byref(self.packed.desc), self.grid_grad_cl.data_ptr(), out.data_ptr(),
                                                          inv.data_ptr(), _stream(self.dev)))
            grads[id(fw.grid)] = out
        return d_film, grads


class RenderFunction(torch.autograd.Function):
    """pixels = render(film, field parameters); see the module docstring."""

    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, film, call, *params):
        module, rd = call['module'], call['rd']
        st = ops.render_forward_stages(module, rd, film, call['x_lin'], call['y_lin'], call['z_lin'], call['cam2world'],
                                       call['rng_perturb'], call['rng_noise_c'], call['rng_u'], call['rng_noise_f'])
        ctx.call, ctx.stages = call, st
        ctx.save_for_backward(film, *params)
        ctx.param_ids = [id(p) for p in call['params']]
        return st['pixels']

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, d_pixels):
        call, st = ctx.call, ctx.stages
        film = ctx.saved_tensors[0]
        module, rd = call['module'], call['rd']
        dev = film.device
        lib = _lib.lib()
        B, n, s = rd.batch, rd.img_h * rd.img_w, rd.num_steps
        c = st['raw_c'].shape[-1]
        hier = bool(rd.hierarchical)
        d_pixels = d_pixels.float().contiguous()
        with torch.cuda.device(dev), torch.no_grad():
            d_raw_c = torch.empty_like(st['raw_c'])
            d_raw_f = torch.empty_like(st['raw_f']) if hier else None
            noise = call['rng_noise_f'] if rd.noise_std != 0.0 else None
            if call.get('grad_rays') is not None:
                # gradient only through the chosen rays: the others were rendered under no_grad in the reference
                mask = torch.zeros(n, dtype=torch.bool, device=dev)
                mask[call['grad_rays']] = True
                d_pixels = d_pixels * mask.reshape(1, 1, rd.img_h, rd.img_w)
            _lib.check(lib.fenerf_composite_backward(
                C.byref(rd), c, st['raw_c'].data_ptr(), st['z_c'].data_ptr(), _ptr(st['raw_f']) if hier else 0,
                _ptr(st['z_f']) if hier else 0, _ptr(noise), d_pixels.data_ptr(), d_raw_c.data_ptr(), _ptr(d_raw_f),
                _stream(dev)))
            m = d_raw_c.abs().max()
            if hier:
                m = torch.maximum(m, d_raw_f.abs().max())
            scale = torch.exp2(4.0 - torch.ceil(torch.log2(m.clamp_min(1e-30)))).float().reshape(1)
            inv_scale = (1.0 / scale).float().reshape(1)
            fb = _FieldBackward(module, film, scale, inv_scale, exact=(rd.precision == _lib.PRECISION['exact']))
            lock = bool(rd.lock_view_dependence)
            rays = call.get('grad_rays')
            dirs = st['dirs']

            def pick(t, last):
                # (B, n, s, last) -> (B, n' * s, last): every ray, or only the rays that carry a gradient (part_forward)
                if rays is not None:
                    t = t.index_select(1, rays)
                return t.reshape(B, -1, last)

            if rays is not None:
                dirs = dirs.index_select(1, rays).contiguous()
            if hier:
                fb.add_points(pick(st['points_f'], 3), dirs, s, lock, pick(st['raw_f'], c), pick(d_raw_f, c))
            fb.add_points(pick(st['points_c'], 3), dirs, s, lock, pick(st['raw_c'], c), pick(d_raw_c, c))
            d_film, grads = fb.finish()
        out = [d_film if ctx.needs_input_grad[0] else None, None]
        for i, pid in enumerate(ctx.param_ids):
            g = grads.get(pid) if ctx.needs_input_grad[2 + i] else None
            if g is not None:
                g = g.reshape(ctx.saved_tensors[1 + i].shape)
            out.append(g)
        return tuple(out)


def render_with_grad(module, rd, film, x_lin, y_lin, z_lin, cam2world, rng_perturb, rng_noise_c, rng_u, rng_noise_f,
                     grad_rays=None):
    """Differentiable render: (B, C-1, R, R) pixels with autograd edges to `film` and the field parameters.
    `grad_rays`: optional int64 ray indices -- only these rays carry the gradient (part_forward)."""
    fw = FieldWeights(module)
    params = fw.parameters()
    call = dict(module=module, rd=rd, x_lin=x_lin, y_lin=y_lin, z_lin=z_lin, cam2world=cam2world, rng_perturb=rng_perturb,
                rng_noise_c=rng_noise_c, rng_u=rng_u, rng_noise_f=rng_noise_f, params=params, grad_rays=grad_rays)
    return RenderFunction.apply(film, call, *params)
