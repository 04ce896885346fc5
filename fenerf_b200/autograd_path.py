"""Torch-op formulation of the render for callers that need autograd -- OPT-IN ONLY.

The fused sm_100a render is forward-only; its backward is SURVEY.md section 8f-1 (not built this
round).  The reference's G step (train_double_latent_semantic.py:411-446) and GAN inversion
(inverse_render_double_semantic.py:385-407) differentiate through the two point-network passes and
the final compositing.  With ``FENERF_B200_TORCH_AUTOGRAD=1`` such calls run here: the
non-differentiable stages (ray set-up, resampling -- ``no_grad`` in the reference too,
generators.py:41, 59) still go through the CUDA library, the differentiable ones are expressed
with torch ops so autograd (and autocast) work.  Without the opt-in a grad-requiring call raises
NotImplementedError: nothing on the render path silently falls back to PyTorch.
"""
import torch
import torch.nn.functional as F

from . import ops
from .generators import volumetric_rendering as vr
from .siren.siren import sample_from_3dgrid


def siren_points_torch(module, points, film, ray_directions):
    """Differentiable restatement of the field (siren/siren.py:164-178, 1509-1530) from the module's
    own nn.Linear layers; `film` is the (B, L, 2, 256) table with 15 f + 30 already applied."""
    spec = module.field_spec()
    x = points * spec.input_scale if spec.input_scale != 1.0 else points
    if ray_directions.shape[1] != points.shape[1]:
        g = points.shape[1] // ray_directions.shape[1]
        ray_directions = ray_directions.unsqueeze(2).expand(-1, -1, g, -1).reshape(points.shape[0], -1, 3)

    def film_layer(layer, h, idx):
        return torch.sin(film[:, idx, 0].unsqueeze(1) * layer.layer(h) + film[:, idx, 1].unsqueeze(1))

    h = x
    for i, layer in enumerate(module.network):
        h = film_layer(layer, h, i)
    sigma = module.final_layer(h)
    parts = [ray_directions]
    if spec.grid_channels:
        parts.append(sample_from_3dgrid(x, module.spatial_embeddings))
    parts.append(h)
    c = torch.cat(parts, dim=-1)
    color = module.color_layer_sine
    color = list(color) if isinstance(color, torch.nn.ModuleList) else [color]
    for j, layer in enumerate(color):
        c = film_layer(layer, c, spec.trunk_layers + j)
    rgb = torch.sigmoid(module.color_layer_linear[0](c))
    outs = [rgb, sigma]
    if spec.label_dim:
        outs.insert(0, module.label_layer_linear(h))
    return torch.cat(outs, dim=-1)


def composite_torch(raw, z_vals, noise, noise_std, clamp_mode, last_back=False, white_back=False, black_back=False):
    """Differentiable alpha compositing over sorted samples (volumetric_rendering.py:18-50)."""
    values, sigmas = raw[..., :-1], raw[..., -1:]
    deltas = torch.cat([z_vals[:, :, 1:] - z_vals[:, :, :-1], 1e10 * torch.ones_like(z_vals[:, :, :1])], -2)
    pre = sigmas + noise * noise_std
    if clamp_mode == 'softplus':
        dens = F.softplus(pre)
    elif clamp_mode == 'relu':
        dens = F.relu(pre)
    else:
        raise TypeError("exceptions must derive from BaseException")  # reference: raise "<str>"
    alphas = 1 - torch.exp(-deltas * dens)
    trans = torch.cumprod(torch.cat([torch.ones_like(alphas[:, :, :1]), 1 - alphas + 1e-10], -2), -2)[:, :, :-1]
    weights = alphas * trans
    weights_sum = weights.sum(2)
    if last_back:
        weights = torch.cat([weights[:, :, :-1], weights[:, :, -1:] + (1 - weights_sum).unsqueeze(2)], 2)
    out = torch.sum(weights * values, -2)
    if white_back:
        out = out + 1 - weights_sum
    if black_back:
        out = out + (1 - weights_sum) * -1
    return out


def _render_autograd(gen, film, batch_size, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev, h_mean,
                     v_mean, hierarchical_sample, sample_dist, lock_view_dependence, kwargs):
    device = torch.device(gen.device)
    rng = kwargs.get('_rng') or vr.DeviceRng(device)
    n_rays = img_size * img_size
    with torch.no_grad():
        rng_perturb = rng.rand(batch_size, n_rays, num_steps, 1).contiguous()
        cam2world, pitch, yaw = ops.camera_poses(batch_size, sample_dist, h_stddev, v_stddev, h_mean, v_mean, rng, device)
        x_lin, y_lin, z_lin = vr.ray_tables(img_size, num_steps, ray_start, ray_end, device)
        rd = ops.make_render_desc(batch=batch_size, img_size=img_size, num_steps=num_steps,
                                  hierarchical=hierarchical_sample, clamp_mode=kwargs['clamp_mode'],
                                  nerf_noise=kwargs['nerf_noise'], fov=fov, precision='exact')
        points, z_vals, dirs, origins = ops.ray_setup(rd, x_lin, y_lin, z_lin, cam2world, rng_perturb)
        dirs_in = dirs
        if lock_view_dependence:
            dirs_in = torch.zeros_like(dirs)
            dirs_in[..., -1] = -1
    coarse = siren_points_torch(gen.siren, points.reshape(batch_size, -1, 3), film, dirs_in)
    coarse = coarse.reshape(batch_size, n_rays, num_steps, -1)
    if hierarchical_sample:
        with torch.no_grad():
            noise_c = rng.randn(batch_size, n_rays, num_steps, 1).contiguous()
            rng_u = rng.rand(batch_size * n_rays, num_steps).contiguous()
            z_fine, points_fine, _ = ops.resample(rd, coarse.detach().float().contiguous(), z_vals, dirs, origins,
                                                  noise_c, rng_u)
        fine = siren_points_torch(gen.siren, points_fine.reshape(batch_size, -1, 3), film, dirs_in)
        fine = fine.reshape(batch_size, n_rays, num_steps, -1)
        all_out = torch.cat([fine, coarse], dim=-2)
        all_z = torch.cat([z_fine, z_vals], dim=-2)
        _, order = torch.sort(all_z, dim=-2)
        all_z = torch.gather(all_z, -2, order)
        all_out = torch.gather(all_out, -2, order.expand(-1, -1, -1, all_out.shape[-1]))
    else:
        all_out, all_z = coarse, z_vals
    noise_f = rng.randn(*all_z.shape)
    pixels = composite_torch(all_out, all_z, noise_f, kwargs['nerf_noise'], kwargs['clamp_mode'],
                             last_back=kwargs.get('last_back', False), white_back=kwargs.get('white_back', False),
                             black_back=kwargs.get('black_back', False))
    if gen.softmax_label:
        pixels = torch.cat([torch.softmax(pixels[..., :-3], dim=-1), pixels[..., -3:]], dim=-1)
    pixels = pixels.reshape(batch_size, img_size, img_size, -1).permute(0, 3, 1, 2).contiguous() * 2 - 1
    return pixels, torch.cat([pitch, yaw], -1)


def _require_opt_in():
    if not ops.autograd_opted_in():
        raise NotImplementedError(ops.GRAD_MESSAGE)


def generator_forward(gen, latents, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev, h_mean, v_mean,
                      hierarchical_sample, sample_dist, lock_view_dependence, kwargs):
    _require_opt_in()
    if len(latents) == 1:
        film = gen.siren.film_table(*gen.siren.mapping_network(latents[0]))
    else:
        f_geo, p_geo = gen.siren.geo_mapping_network(latents[0])
        f_app, p_app = gen.siren.app_mapping_network(latents[1])
        film = gen.siren.film_table(f_geo, f_app, p_geo, p_app)
    return _render_autograd(gen, film, latents[0].shape[0], img_size, fov, ray_start, ray_end, num_steps, h_stddev,
                            v_stddev, h_mean, v_mean, hierarchical_sample, sample_dist, lock_view_dependence, kwargs)


def generator_forward_with_frequencies(gen, film_inputs, img_size, fov, ray_start, ray_end, num_steps, h_stddev,
                                       v_stddev, h_mean, v_mean, hierarchical_sample, sample_dist,
                                       lock_view_dependence, kwargs):
    _require_opt_in()
    film = gen.siren.film_table(*film_inputs)
    return _render_autograd(gen, film, film_inputs[0].shape[0], img_size, fov, ray_start, ray_end, num_steps, h_stddev,
                            v_stddev, h_mean, v_mean, hierarchical_sample, sample_dist, lock_view_dependence, kwargs)
