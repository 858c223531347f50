"""Oracle for the image-space loss that directly follows the render call.

TEST INFRASTRUCTURE ONLY.  Restates, in plain torch:

* the blend + clamp of ``freegaussian_model.py:875-877``: ``rgb = clamp(render[..., :3] + (1 - alpha) * background, 0, 1)``;
* ``get_loss_dict`` (``freegaussian_model.py:965-981``): ``(1 - l) * |gt - pred|.mean() + l * (1 - SSIM(gt, pred))``
  with ``ssim_lambda = 0.2``;
* ``pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)`` (``freegaussian_model.py:22, 217``), an
  un-vendored dependency that is not installed here: 11-tap Gaussian window, sigma 1.5, separable "valid"
  convolution (H first, then W), K = (0.01, 0.03), mean over the valid region, channels and batch.
  PARITY UNPINNED for the SSIM half (published algorithm restated; cross-checked against the window-by-window
  definition in tests/test_loss.py); the L1 / blend half is the reference's own code.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def gaussian_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32, device=None) -> Tensor:
    coords = torch.arange(size, dtype=dtype, device=device) - size // 2
    g = torch.exp(-(coords**2) / (2 * sigma**2))
    return g / g.sum()


def _filter(x: Tensor, win: Tensor) -> Tensor:
    C = x.shape[1]
    k = win.numel()
    out = x
    if out.shape[2] >= k:
        out = F.conv2d(out, win.view(1, 1, k, 1).repeat(C, 1, 1, 1), groups=C)
    if out.shape[3] >= k:
        out = F.conv2d(out, win.view(1, 1, 1, k).repeat(C, 1, 1, 1), groups=C)
    return out


def ssim(X: Tensor, Y: Tensor, data_range: float = 1.0, K=(0.01, 0.03)) -> Tensor:
    """X, Y [B,C,H,W] -> scalar mean SSIM."""
    win = gaussian_window(dtype=X.dtype, device=X.device)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(X, win), _filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = _filter(X * X, win) - mu1_sq
    sigma2_sq = _filter(Y * Y, win) - mu2_sq
    sigma12 = _filter(X * Y, win) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return ssim_map.flatten(2).mean(-1).mean()


def blend_l1_ssim_loss(render: Tensor, alpha: Tensor, background: Tensor, gt: Tensor, ssim_lambda: float = 0.2,
                       mask: Tensor = None) -> Tensor:
    """render [H,W,>=3] (premultiplied, first three channels RGB), alpha [H,W,1], background [3], gt [H,W,3],
    mask [H,W,1] (optional)."""
    pred = torch.clamp(render[..., :3] + (1 - alpha) * background, 0.0, 1.0)  # model.py:876-877
    if mask is not None:  # model.py:957-963
        gt = gt * mask
        pred = pred * mask
    l1 = torch.abs(gt - pred).mean()  # :965
    sim = 1 - ssim(gt.permute(2, 0, 1)[None], pred.permute(2, 0, 1)[None])  # :966
    return (1 - ssim_lambda) * l1 + ssim_lambda * sim  # :981


def depth_fixup(render: Tensor, alpha: Tensor) -> Tensor:
    """``freegaussian_model.py:884-886`` verbatim (before the ``squeeze(0)``): render [1,H,W,4] "RGB+ED", alpha [1,H,W,1]."""
    depth_im = render[:, ..., 3:4]
    return torch.where(alpha > 0, depth_im, depth_im.detach().max())
