"""voltools_b200 -- B200-native drop-in for the GPU resampling path of the-lay/voltools (v0.6.0 API)."""
__version__ = '0.6.0+b200.1'

from .transforms import AVAILABLE_INTERPOLATIONS, AVAILABLE_DEVICES, scale, shear, rotate, translate, transform, affine, project, \
    release_host_buffers
from .volume import StaticVolume
from ._native import pinned_empty
from . import utils
