from . import generators, volumetric_rendering  # noqa: F401
from .generators import ImplicitGenerator3d, DoubleImplicitGenerator3d  # noqa: F401
