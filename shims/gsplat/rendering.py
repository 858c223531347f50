"""``from gsplat.rendering import rasterization`` (``freegaussian/freegaussian_model.py:18``)."""
from freegaussian_b200.rendering import rasterization, rasterize_to_pixels  # noqa: F401
