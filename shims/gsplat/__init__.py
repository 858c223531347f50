"""``gsplat`` import shim: put ``<repo>/shims`` (and the repo root) on ``PYTHONPATH`` and the reference's import lines
(``freegaussian/freegaussian_model.py:15-21``, ``freegaussian_control_model.py:7-10``,
``preprocess/knn_gaussian.py:9``, ``render_color.py:9``, ``render_depth.py:9``, ``o3d_color_splat.py:11``) resolve to
freegaussian_b200 without touching the reference's sources (first drop-in option).  Only the three
modules the reference imports exist; nothing else of gsplat's surface is claimed."""
from freegaussian_b200.rendering import rasterization  # noqa: F401

__version__ = "1.4.0+freegaussian_b200"  # the gsplat semantics the renderer follows (2-D radii, cuda_legacy present)
