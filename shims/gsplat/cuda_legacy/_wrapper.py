"""``from gsplat.cuda_legacy._wrapper import num_sh_bases`` (``freegaussian/freegaussian_model.py:21``, used ``:165``)."""
from freegaussian_b200.compat import num_sh_bases  # noqa: F401
