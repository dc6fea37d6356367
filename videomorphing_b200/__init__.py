"""videomorphing_b200 -- B200-native (sm_100a) halfway-domain optimizer + morph renderer of liaojing/videomorphing.

The package is a thin host-side mirror of the reference's Algorithm/ operator surface over the C-ABI CUDA library
libvmorph.so (include/vmorph.h).  Importing it does not need a GPU; every compute call does.
"""
from . import _lib  # noqa: F401
from .api import (BCOND_BORDER, BCOND_CORNER, BCOND_NONE, REFERENCE_VOXEL_CAP, Morph, Parameters, Pyramid,  # noqa: F401
                  level_schedule, quadratic_path, render_halfway_image, render_sequence, stencils)
