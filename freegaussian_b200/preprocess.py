"""GPU version of the attribute-mask assignment loop body of ``preprocess/knn_gaussian.py:116-132``
(the other caller of the ``rasterization`` boundary: ``packed=True``, ``"ED"``).

    render, alpha, info = rasterization(..., packed=True, render_mode="ED", sh_degree=3)   # :93-113
    assign_gaussian_masks(render, info, mask, gaussian_masks)                                # :116-132

``mask`` is the reference's ``data["atrb_masks"][..., :-1] & data["mask_valids"][..., :-1]`` (``:128``),
``gaussian_masks`` its ``[N, M]`` bool accumulator (``:58``), updated in place on the GPU.
"""

from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr


@torch.no_grad()
@_lib.on_device_of("render")
def assign_gaussian_masks(render: Tensor, info: dict, atrb_masks: Tensor, gaussian_masks: Tensor,
                          mask_valids: Tensor | None = None) -> Tensor:
    """render [1,H,W,1] expected depth; info: packed meta (means2d [nnz,2], depths [nnz], gaussian_ids [nnz]);
    atrb_masks [H,W,M] bool; mask_valids [M] bool (None = all valid); gaussian_masks [N,M] bool (in place)."""
    depth = render.squeeze()
    H, W = depth.shape
    M = atrb_masks.shape[-1]
    assert atrb_masks.shape == (H, W, M) and gaussian_masks.shape[1] == M
    for t in (depth, atrb_masks, gaussian_masks, info["means2d"]):
        if not t.is_cuda:
            raise RuntimeError("assign_gaussian_masks has no CPU path")
    assert gaussian_masks.dtype == torch.bool and gaussian_masks.is_contiguous()
    valid = torch.ones(M, dtype=torch.bool, device=depth.device) if mask_valids is None else mask_valids.to(depth.device)
    means2d = info["means2d"].detach().contiguous()
    depths = info["depths"].detach().contiguous()
    gids = info["gaussian_ids"].to(torch.int64).contiguous()
    check(_lib.lib().fg_assign_masks(
        means2d.shape[0], ptr(means2d), ptr(depths), ptr(gids), ptr(depth.contiguous()), W, H,
        ptr(atrb_masks.contiguous().view(torch.uint8)), ptr(valid.contiguous().view(torch.uint8)), M,
        ptr(gaussian_masks.view(torch.uint8)), torch.cuda.current_stream().cuda_stream))
    return gaussian_masks
