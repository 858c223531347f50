"""``k_nearest`` -- GPU replacement for ``FreeGaussianModel.k_nearest_sklearn``
(``freegaussian/freegaussian_model.py:293-311``, used at ``:158-162`` to seed the scales).

Same contract: exact Euclidean k-NN of the point set against itself, k+1 neighbours with
the self column dropped; distances bit-identical to sklearn's (float64 arithmetic,
returned as float32).  Indices are returned as int32 rather than the reference's lossy
float32 cast (``:311``; the reference discards them, ``:158``).
"""

from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr


@_lib.on_device_of("x")
def k_nearest(x: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """x [N,3] float32 CUDA -> (distances [N,k] float32, indices [N,k] int32), ascending."""
    assert x.dim() == 2 and x.shape[1] == 3, x.shape
    assert x.dtype == torch.float32, x.dtype
    if not x.is_cuda:
        raise RuntimeError("freegaussian_b200.k_nearest has no CPU path; move the points to the GPU")
    L = _lib.lib()
    n = x.shape[0]
    assert n > k >= 1, (n, k)
    x = x.contiguous()
    dist = torch.empty(n, k, dtype=torch.float32, device=x.device)
    idx = torch.empty(n, k, dtype=torch.int32, device=x.device)
    ws = torch.empty(L.fg_knn_workspace_bytes(n), dtype=torch.uint8, device=x.device)
    check(L.fg_knn_f32(n, ptr(x), k, ptr(dist), ptr(idx), ptr(ws), ws.numel(), torch.cuda.current_stream().cuda_stream))
    return dist, idx
