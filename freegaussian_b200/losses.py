"""Fused image-space loss that directly follows the render call.

``blend_l1_ssim_loss(render, alpha, background, gt)`` equals, in one forward and one backward kernel,

    rgb  = clamp(render[..., :3] + (1 - alpha) * background, 0, 1)            # freegaussian_model.py:876-877
    loss = (1 - l) * |gt - rgb|.mean() + l * (1 - SSIM(gt, rgb))              # freegaussian_model.py:965-981

with ``SSIM = pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)`` (``:217``) and
``l = ssim_lambda = 0.2``.  Gradients flow to ``render`` and ``alpha``.
"""

from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr


class _BlendL1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, alpha, background, gt, ssim_lambda):
        L = _lib.lib()
        H, W, rs = render.shape[-3], render.shape[-2], render.shape[-1]
        render_c, alpha_c = render.contiguous(), alpha.contiguous()
        bg, gt_c = background.contiguous().float(), gt.contiguous()
        dev = render.device
        partial = torch.empty(L.fg_l1_ssim_workspace_floats(W, H), device=dev)
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        check(L.fg_l1_ssim_fwd(W, H, rs, ptr(render_c), ptr(alpha_c), ptr(bg), ptr(gt_c), float(ssim_lambda),
                               ptr(partial), ptr(sums), st))
        ctx.save_for_backward(render_c, alpha_c, bg, gt_c, partial)
        ctx.meta = (W, H, rs, float(ssim_lambda), render.shape, alpha.shape)
        # loss = L1 term + lambda * (1 - mean SSIM); sums[1] already holds lambda * mean SSIM
        return (sums[0] + ssim_lambda - sums[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, v_loss):
        L = _lib.lib()
        render, alpha, bg, gt, partial = ctx.saved_tensors
        W, H, rs, lam, rshape, ashape = ctx.meta
        v_render = torch.empty_like(render)
        v_alpha = torch.empty_like(alpha)
        vl = v_loss.reshape(1).to(torch.float32).contiguous()
        check(L.fg_l1_ssim_bwd(W, H, rs, ptr(render), ptr(alpha), ptr(bg), ptr(gt), lam, ptr(partial), ptr(vl),
                               ptr(v_render), ptr(v_alpha), torch.cuda.current_stream().cuda_stream))
        return v_render.view(rshape), v_alpha.view(ashape), None, None, None


def blend_l1_ssim_loss(render: Tensor, alpha: Tensor, background: Tensor, gt: Tensor, ssim_lambda: float = 0.2) -> Tensor:
    """render [1,H,W,>=3] or [H,W,>=3] (premultiplied RGB first), alpha [..,H,W,1], background [3], gt [H,W,3]."""
    for name, t in (("render", render), ("alpha", alpha), ("background", background), ("gt", gt)):
        if not t.is_cuda:
            raise RuntimeError(f"blend_l1_ssim_loss: `{name}` is not a CUDA tensor (no CPU path)")
    assert render.shape[-3:-1] == gt.shape[-3:-1] and gt.shape[-1] == 3 and render.shape[-1] >= 3
    assert render.numel() == render.shape[-3] * render.shape[-2] * render.shape[-1], "one image per call"
    return _BlendL1SSIM.apply(render, alpha, background, gt, ssim_lambda)
