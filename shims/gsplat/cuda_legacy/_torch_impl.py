"""``from gsplat.cuda_legacy._torch_impl import quat_to_rotmat`` (``freegaussian/freegaussian_model.py:15``, used ``:535``)."""
from freegaussian_b200.compat import quat_to_rotmat  # noqa: F401
