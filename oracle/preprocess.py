"""Oracle for the attribute-mask assignment of ``preprocess/knn_gaussian.py:116-132``.

TEST INFRASTRUCTURE ONLY.  PARITY PINNED: the logic is a few lines of plain torch indexing that live in
the reference repository itself, restated here statement by statement (CPU tensors).
"""

from __future__ import annotations

import torch
from torch import Tensor


def assign_gaussian_masks_reference(render: Tensor, means2d: Tensor, depths: Tensor, gaussian_ids: Tensor,
                                    mask: Tensor, gaussian_masks: Tensor) -> Tensor:
    """render [1,H,W,1]; means2d [nnz,2]; depths [nnz]; gaussian_ids [nnz]; mask [H,W,M] bool (already combined with
    mask_valids, knn_gaussian.py:128); gaussian_masks [N,M] bool, updated in place like :132."""
    depth = render.squeeze()  # :119
    H, W = depth.shape
    M = mask.shape[-1]
    xy = means2d.cpu().long()  # :116
    im = ((xy >= 0) & (xy < torch.tensor([W, H]))).all(-1)  # :117
    xy = xy[im]  # :118
    d_at = depth[xy[:, 1], xy[:, 0]]
    delta_depth = d_at - depths[im]  # :120
    dm = (-d_at * 0.1 < delta_depth) & (delta_depth < d_at * 1)  # :121
    xy = xy[dm]  # :123
    m = mask[xy[:, 1], xy[:, 0]]  # :129
    ids = gaussian_ids[im][dm]  # :130
    ids = ids[..., None].expand(-1, M)[m]  # :131
    gaussian_masks[ids, m.nonzero()[:, -1]] = True  # :132
    return gaussian_masks
