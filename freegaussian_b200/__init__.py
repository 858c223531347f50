"""freegaussian_b200 -- B200-native (sm_100a) splat renderer behind FreeGaussian's
``rasterization(...)`` call (``freegaussian/freegaussian_model.py:847-868``).

Drop-in use (the reference imports ``from gsplat.rendering import rasterization``,
``freegaussian_model.py:17-20``)::

    from freegaussian_b200.rendering import rasterization
    from freegaussian_b200.compat import num_sh_bases, quat_to_rotmat

Only host logic lives in Python; all arithmetic runs in ``libfreegaussian_b200.so``
(``include/fg_api.h``).  Importing this package does not load the library; the first
call does, and raises if it has not been built.
"""

from .compat import get_viewmat, num_sh_bases, quat_to_rotmat  # noqa: F401
from .knn import k_nearest  # noqa: F401
from .rendering import isect_tiles, rasterization, rasterize_to_pixels  # noqa: F401
from .optim import GaussianAdam  # noqa: F401
from .densify import refine  # noqa: F401

__all__ = [
    "rasterization", "rasterize_to_pixels", "isect_tiles", "k_nearest", "num_sh_bases", "quat_to_rotmat",
    "get_viewmat", "GaussianAdam", "refine",
]
