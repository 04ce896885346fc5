"""fenerf_b200 -- B200-native volumetric face renderer behind FENeRF's generator API.

Layout (only what the hot path needs):
  csrc/                 sm_100a CUDA kernels + the C-ABI (include/fenerf_b200.h)
  _lib.py, ops.py       ctypes binding and thin operators (plumbing)
  packing.py            raw nn.Parameter pointers -> fenerf_pack_field
  generators/, siren/   host-side mirror of the reference's class API (drop-in boundary)
  dist.py               batch sharding over ranks + the NCCL frame all-gather
"""
import sys

from . import _lib, ops  # noqa: F401
from .generators import generators as _generators_mod
from .generators import volumetric_rendering as _vr_mod
from .siren import siren as _siren_mod

__all__ = ["install", "ops", "generators", "siren"]


def install():
    """Make ``import generators.generators`` / ``import siren.siren`` resolve to this package.

    The reference's scripts do ``import generators`` / ``import siren`` and look classes up with
    ``getattr(generators, metadata['generator'])`` (train_double_latent_semantic.py:142); whole-module
    checkpoints are pickled as ``generators.generators.<Class>`` / ``siren.siren.<Class>``
    (render_multiview_images_double_semantic.py:58).  Call this once before those imports (or put
    it in sitecustomize) and the scripts run unchanged on top of the B200 library.
    """
    from . import generators as gen_pkg
    from . import siren as siren_pkg
    sys.modules['generators'] = gen_pkg
    sys.modules['generators.generators'] = _generators_mod
    sys.modules['generators.volumetric_rendering'] = _vr_mod
    sys.modules['siren'] = siren_pkg
    sys.modules['siren.siren'] = _siren_mod
    # the reference's `generators/__init__.py` does `from .generators import *`-style exposure via
    # getattr(generators, name): mirror the two class names at package level
    for name in ("ImplicitGenerator3d", "DoubleImplicitGenerator3d"):
        setattr(gen_pkg, name, getattr(_generators_mod, name))
    for name in dir(_siren_mod):
        if not name.startswith('_'):
            setattr(siren_pkg, name, getattr(_siren_mod, name))
    return gen_pkg, siren_pkg
