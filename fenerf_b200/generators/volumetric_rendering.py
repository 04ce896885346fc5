"""Host-side pieces of the render path that stay in PyTorch, plus the RNG draw protocol.

What lives here is what SURVEY.md section 8a marks "~0 % of the time" and section 7 (hard part 5)
says to keep in torch for ulp-compatibility with the reference: camera pose sampling (a3), the 4x4
look-at matrix (a4) and the three tiny linspace tables.  Everything per-ray or per-sample is in the
CUDA library (csrc/rays.cu, resample.cu, composite.cu).

Reference functions mirrored (generators/volumetric_rendering.py):
  sample_camera_positions :179-228   create_cam2world_matrix :230-248   truncated_normal_ :170-177
  normalize_vecs (generators/math_utils_torch.py:16-20)
"""
import math
import random

import torch


# --------------------------------------------------------------------------------------------
# RNG protocol
# --------------------------------------------------------------------------------------------
class DeviceRng:
    """Draws on ``device`` from torch's global generator, in the order the reference draws:
    #1 rand(B,N,S,1) perturb -> camera draws -> #4 randn(B,N,S,1) -> #5 rand(B*N,S) -> #6 randn(B,N,S',1)
    (SURVEY.md section 8a).  The noise tensors are drawn even when nerf_noise == 0 because the
    reference does (volumetric_rendering.py:27), which keeps the generator stream in step."""

    def __init__(self, device):
        self.device = device

    def rand(self, *shape):
        return torch.rand(shape, device=self.device)

    def randn(self, *shape):
        return torch.randn(shape, device=self.device)

    def coin(self):
        return random.random()


class ReplayRng:
    """Replays draws recorded from an oracle run (tests): a list of (kind, tensor)."""

    def __init__(self, draws, device):
        self.draws = list(draws)
        self.device = device
        self.pos = 0

    def _next(self, kind, shape):
        if self.pos >= len(self.draws):
            raise RuntimeError("ReplayRng exhausted at draw %d (%s %s)" % (self.pos, kind, shape))
        k, t = self.draws[self.pos]
        self.pos += 1
        if k != kind or tuple(t.shape) != tuple(shape):
            raise RuntimeError("ReplayRng mismatch at draw %d: recorded %s%s, requested %s%s"
                               % (self.pos - 1, k, tuple(t.shape), kind, tuple(shape)))
        return t.to(self.device)

    def rand(self, *shape):
        return self._next("rand", shape)

    def randn(self, *shape):
        return self._next("randn", shape)

    def coin(self):
        k, t = self.draws[self.pos]
        self.pos += 1
        assert k == "coin"
        return float(t)


# --------------------------------------------------------------------------------------------
# camera
# --------------------------------------------------------------------------------------------
def normalize_vecs(vectors):
    return vectors / (torch.norm(vectors, dim=-1, keepdim=True))


def truncated_normal_(tensor, mean=0, std=1, rng=None):
    """Resample-free truncation to (-2, 2): first of four normal draws that lands inside."""
    size = tensor.shape
    tmp = tensor.new_empty(size + (4,)).normal_() if rng is None else rng.randn(*size, 4)
    valid = (tmp < 2) & (tmp > -2)
    ind = valid.max(-1, keepdim=True)[1]
    tensor.data.copy_(tmp.gather(-1, ind).squeeze(-1))
    tensor.data.mul_(std).add_(mean)
    return tensor


def sample_camera_positions(device, n=1, r=1, horizontal_stddev=1, vertical_stddev=1, horizontal_mean=math.pi * 0.5,
                            vertical_mean=math.pi * 0.5, mode='normal', rng=None):
    """n camera origins on the radius-r sphere; theta = yaw, phi = pitch (clamped to (1e-5, pi-1e-5)).

    Same distributions and the same draw order (theta before phi) as the reference."""
    rng = rng or DeviceRng(device)

    def uniform(stddev, mean, widen=1):
        return (rng.rand(n, 1) - 0.5) * 2 * stddev * widen + mean

    def gaussian(stddev, mean):
        return rng.randn(n, 1) * stddev + mean

    if mode == 'uniform':
        theta = uniform(horizontal_stddev, horizontal_mean)
        phi = uniform(vertical_stddev, vertical_mean)
    elif mode == 'normal' or mode == 'gaussian':
        theta = gaussian(horizontal_stddev, horizontal_mean)
        phi = gaussian(vertical_stddev, vertical_mean)
    elif mode == 'hybrid':
        if rng.coin() < 0.5:
            theta = (rng.rand(n, 1) - 0.5) * 2 * horizontal_stddev * 2 + horizontal_mean
            phi = (rng.rand(n, 1) - 0.5) * 2 * vertical_stddev * 2 + vertical_mean
        else:
            theta = gaussian(horizontal_stddev, horizontal_mean)
            phi = gaussian(vertical_stddev, vertical_mean)
    elif mode == 'truncated_gaussian':
        theta = truncated_normal_(torch.zeros((n, 1), device=device), rng=rng) * horizontal_stddev + horizontal_mean
        phi = truncated_normal_(torch.zeros((n, 1), device=device), rng=rng) * vertical_stddev + vertical_mean
    elif mode == 'spherical_uniform':
        theta = (rng.rand(n, 1) - .5) * 2 * horizontal_stddev + horizontal_mean
        v_stddev, v_mean = vertical_stddev / math.pi, vertical_mean / math.pi
        v = ((rng.rand(n, 1) - .5) * 2 * v_stddev + v_mean)
        v = torch.clamp(v, 1e-5, 1 - 1e-5)
        phi = torch.arccos(1 - 2 * v)
    else:
        theta = torch.ones((n, 1), device=device, dtype=torch.float) * horizontal_mean
        phi = torch.ones((n, 1), device=device, dtype=torch.float) * vertical_mean

    phi = torch.clamp(phi, 1e-5, math.pi - 1e-5)
    origins = torch.zeros((n, 3), device=device)
    origins[:, 0:1] = r * torch.sin(phi) * torch.cos(theta)
    origins[:, 2:3] = r * torch.sin(phi) * torch.sin(theta)
    origins[:, 1:2] = r * torch.cos(phi)
    return origins, phi, theta


def create_cam2world_matrix(forward_vector, origin, device=None):
    """Look-at camera-to-world: R = [-left, up, -forward] columns, then translate to origin."""
    forward_vector = normalize_vecs(forward_vector)
    up = torch.tensor([0, 1, 0], dtype=torch.float, device=device).expand_as(forward_vector)
    left = normalize_vecs(torch.cross(up, forward_vector, dim=-1))
    up = normalize_vecs(torch.cross(forward_vector, left, dim=-1))
    n = forward_vector.shape[0]
    rotation = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    rotation[:, :3, :3] = torch.stack((-left, up, -forward_vector), axis=-1)
    translation = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    translation[:, :3, 3] = origin
    return translation @ rotation


def ray_tables(img_size, num_steps, ray_start, ray_end, device):
    """The three linspace tables get_initial_rays_trig builds (volumetric_rendering.py:115-124);
    torch.linspace is kept so the values are the reference's to the bit."""
    x_lin = torch.linspace(-1, 1, img_size, device=device)
    y_lin = torch.linspace(1, -1, img_size, device=device)
    z_lin = torch.linspace(ray_start, ray_end, num_steps, device=device)
    return x_lin, y_lin, z_lin
