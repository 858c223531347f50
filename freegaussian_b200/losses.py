"""Fused image-space loss that directly follows the render call.

``blend_l1_ssim_loss(render, alpha, background, gt)`` equals, in one forward and one backward kernel,

    rgb  = clamp(render[..., :3] + (1 - alpha) * background, 0, 1)            # freegaussian_model.py:876-877
    loss = (1 - l) * |gt - rgb|.mean() + l * (1 - SSIM(gt, rgb))              # freegaussian_model.py:965-981

with ``SSIM = pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)`` (``:217``) and
``l = ssim_lambda = 0.2``.  Gradients flow to ``render`` and ``alpha``.  ``mask`` (``[H,W,1]``) multiplies both images
first, as ``get_loss_dict`` does for masked batches (``:957-963``).

``depth_fixup(render, alpha)`` is the other post-render step of ``get_outputs`` (``:884-886``):
``depth = where(alpha > 0, ED, ED.detach().max())``.
"""

from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr


class _BlendL1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, alpha, background, gt, ssim_lambda, mask=None):
        L = _lib.lib()
        H, W, rs = render.shape[-3], render.shape[-2], render.shape[-1]
        render_c, alpha_c = render.contiguous(), alpha.contiguous()
        bg, gt_c = background.contiguous().float(), gt.contiguous()
        mask_c = None if mask is None else mask.to(torch.float32).contiguous()
        dev = render.device
        partial = torch.empty(L.fg_l1_ssim_workspace_floats(W, H), device=dev)
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        check(L.fg_l1_ssim_fwd(W, H, rs, ptr(render_c), ptr(alpha_c), ptr(bg), ptr(gt_c), ptr(mask_c), float(ssim_lambda),
                               ptr(partial), ptr(sums), st))
        ctx.save_for_backward(render_c, alpha_c, bg, gt_c, partial, mask_c)
        ctx.meta = (W, H, rs, float(ssim_lambda), render.shape, alpha.shape)
        # loss = L1 term + lambda * (1 - mean SSIM); sums[1] already holds lambda * mean SSIM
        return (sums[0] + ssim_lambda - sums[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, v_loss):
        L = _lib.lib()
        render, alpha, bg, gt, partial, mask = ctx.saved_tensors
        W, H, rs, lam, rshape, ashape = ctx.meta
        v_render = torch.empty_like(render)
        v_alpha = torch.empty_like(alpha)
        vl = v_loss.reshape(1).to(torch.float32).contiguous()
        check(L.fg_l1_ssim_bwd(W, H, rs, ptr(render), ptr(alpha), ptr(bg), ptr(gt), ptr(mask), lam, ptr(partial), ptr(vl),
                               ptr(v_render), ptr(v_alpha), torch.cuda.current_stream().cuda_stream))
        return v_render.view(rshape), v_alpha.view(ashape), None, None, None, None


@_lib.on_device_of("render")
def blend_l1_ssim_loss(render: Tensor, alpha: Tensor, background: Tensor, gt: Tensor, ssim_lambda: float = 0.2,
                       mask: Tensor = None) -> Tensor:
    """render [1,H,W,>=3] or [H,W,>=3] (premultiplied RGB first), alpha [..,H,W,1], background [3], gt [H,W,3],
    mask [H,W,1] or [H,W] (optional; bool or float)."""
    for name, t in (("render", render), ("alpha", alpha), ("background", background), ("gt", gt)):
        if not t.is_cuda:
            raise RuntimeError(f"blend_l1_ssim_loss: `{name}` is not a CUDA tensor (no CPU path)")
    assert render.shape[-3:-1] == gt.shape[-3:-1] and gt.shape[-1] == 3 and render.shape[-1] >= 3
    assert render.numel() == render.shape[-3] * render.shape[-2] * render.shape[-1], "one image per call"
    if mask is not None:
        assert mask.is_cuda and mask.numel() == gt.shape[-3] * gt.shape[-2], mask.shape
    return _BlendL1SSIM.apply(render, alpha, background, gt, ssim_lambda, mask)


class _DepthFixup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, alpha, channel):
        L = _lib.lib()
        render_c, alpha_c = render.contiguous(), alpha.contiguous()
        stride = render.shape[-1]
        n = render_c.numel() // stride
        depth = torch.empty(alpha.shape, device=render.device)
        ws = torch.empty(1, dtype=torch.int32, device=render.device)
        check(L.fg_depth_fixup_fwd(n, ptr(render_c), stride, channel, ptr(alpha_c), ptr(depth), ptr(ws),
                                   torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(alpha_c)
        ctx.meta = (n, stride, channel, render.shape)
        return depth

    @staticmethod
    def backward(ctx, v_depth):
        (alpha,) = ctx.saved_tensors
        n, stride, channel, rshape = ctx.meta
        v_render = torch.empty(rshape, device=alpha.device)
        check(_lib.lib().fg_depth_fixup_bwd(n, ptr(alpha), ptr(v_depth.contiguous()), stride, channel, ptr(v_render),
                                            torch.cuda.current_stream().cuda_stream))
        return v_render, None, None


@_lib.on_device_of("render")
def depth_fixup(render: Tensor, alpha: Tensor, channel: int = 3) -> Tensor:
    """``torch.where(alpha > 0, render[..., 3:4], render[..., 3:4].detach().max())`` (``freegaussian_model.py:884-886``) for
    an "RGB+ED" render ``[C,H,W,4]`` / alpha ``[C,H,W,1]``; returns ``[C,H,W,1]`` (the reference then squeezes dim 0).
    The maximum runs over the whole tensor, like the reference's."""
    if not (render.is_cuda and alpha.is_cuda):
        raise RuntimeError("depth_fixup: inputs are not CUDA tensors (no CPU path)")
    assert render.dtype == torch.float32 and alpha.shape[:-1] == render.shape[:-1] and alpha.shape[-1] == 1
    assert 0 <= channel < render.shape[-1]
    return _DepthFixup.apply(render, alpha, channel)
