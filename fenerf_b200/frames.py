"""Frame consumers on the GPU: drop-ins for the two per-frame CPU loops of the
reference's render / sampling / FID scripts.

  mask2color(masks)        train_double_latent_semantic.py:66-72 (imported by every render_* script)
  frames_to_uint8(frames)  the normalize + quantise step of torchvision.utils.save_image (fid_evaluation.py:146-151)
"""
import torch

from . import _lib


def _dev(t):
    if t.is_cuda:
        return t, None
    if not torch.cuda.is_available():
        raise RuntimeError("fenerf_b200.frames needs a CUDA device")
    return t.cuda(non_blocking=True), t.device


def mask2color(masks):
    """(B, K, H, W) label scores -> (B, 3, H, W) float colours 0..255 (the reference's COLOR_MAP).  A CPU input
    (what staged_forward returns) is coloured on the GPU and handed back on the CPU, as the reference returns it."""
    m, back = _dev(masks)
    m = m.float().contiguous()
    b, k, h, w = m.shape
    out = torch.empty((b, 3, h, w), dtype=torch.float32, device=m.device)
    with torch.cuda.device(m.device):
        _lib.check(_lib.lib().fenerf_mask2color(m.data_ptr(), b, k, h * w, out.data_ptr(),
                                                torch.cuda.current_stream(m.device).cuda_stream))
    return out if back is None else out.to(back)


def frames_to_uint8(frames, channels=None):
    """(B, C, H, W) frames in [-1, 1] -> (B, H, W, n) uint8 (``channels`` = (first, n); default the last three = rgb)."""
    f, back = _dev(frames)
    f = f.float().contiguous()
    b, c, h, w = f.shape
    c0, nc = channels if channels is not None else (c - 3, 3)
    out = torch.empty((b, h, w, nc), dtype=torch.uint8, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(_lib.lib().fenerf_frames_to_u8(f.data_ptr(), b, c, c0, nc, h * w, out.data_ptr(),
                                                  torch.cuda.current_stream(f.device).cuda_stream))
    return out if back is None else out.to(back)
