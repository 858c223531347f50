"""Fused blend + L1 + SSIM loss: oracle self-checks on CPU, kernel parity on GPU."""
import pytest
import torch

from oracle import loss as OL
from util import grad_rel_err


def _inputs(H, W, seed, rs=4):
    g = torch.Generator().manual_seed(seed)
    alpha = torch.rand(1, H, W, 1, generator=g)
    render = torch.rand(1, H, W, rs, generator=g) * alpha * 1.3 - 0.05  # some pixels clamp at both ends
    bg = torch.tensor([0.149, 0.1647, 0.2157])  # the reference's default background (model.py:223)
    gt = torch.rand(H, W, 3, generator=g)
    return render, alpha, bg, gt


def test_oracle_ssim_properties():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 40, 52, generator=g)
    assert abs(float(OL.ssim(x, x)) - 1.0) < 1e-6  # identical images
    y = torch.rand(1, 3, 40, 52, generator=g)
    assert abs(float(OL.ssim(x, y)) - float(OL.ssim(y, x))) < 1e-6  # symmetric
    assert float(OL.ssim(x, y)) < 0.2
    w = OL.gaussian_window()
    assert abs(float(w.sum()) - 1) < 1e-6 and w.argmax() == 5


def test_oracle_ssim_matches_the_windowed_definition():
    """SSIM as defined by Wang et al. with the 11x11 Gaussian window (sigma 1.5) evaluated window by window over the valid
    region in plain float64 loops, against the oracle's separable-convolution form (what pytorch_msssim implements)."""
    import math

    import numpy as np

    g = torch.Generator().manual_seed(4)
    X, Y = torch.rand(1, 2, 15, 18, generator=g).double(), torch.rand(1, 2, 15, 18, generator=g).double()
    w1 = np.array([math.exp(-((i - 5) ** 2) / (2 * 1.5 ** 2)) for i in range(11)])
    w1 /= w1.sum()
    w2 = np.outer(w1, w1)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    vals = []
    for c in range(2):
        x, y = X[0, c].numpy(), Y[0, c].numpy()
        for i in range(15 - 10):
            for j in range(18 - 10):
                px, py = x[i:i + 11, j:j + 11], y[i:i + 11, j:j + 11]
                mx, my = (w2 * px).sum(), (w2 * py).sum()
                vx, vy, cxy = (w2 * px * px).sum() - mx * mx, (w2 * py * py).sum() - my * my, (w2 * px * py).sum() - mx * my
                vals.append((2 * mx * my + C1) * (2 * cxy + C2) / ((mx * mx + my * my + C1) * (vx + vy + C2)))
    assert abs(float(OL.ssim(X, Y)) - float(np.mean(vals))) < 1e-12


def test_oracle_loss_gradcheck():
    render, alpha, bg, gt = _inputs(14, 15, 1)
    r, a = render.double().requires_grad_(True), alpha.double().requires_grad_(True)
    f = lambda r_, a_: OL.blend_l1_ssim_loss(r_[0], a_[0], bg.double(), gt.double())
    assert torch.autograd.gradcheck(f, (r, a), eps=1e-6, atol=1e-5, nondet_tol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,rs", [(11, 11, 3), (37, 53, 4), (128, 200, 4), (540, 960, 4)])
def test_fused_loss_matches_oracle(built_lib, H, W, rs):
    from freegaussian_b200.losses import blend_l1_ssim_loss
    render, alpha, bg, gt = _inputs(H, W, H + W, rs)
    ro, ao = render.clone().double().requires_grad_(True), alpha.clone().double().requires_grad_(True)
    ref = OL.blend_l1_ssim_loss(ro[0], ao[0], bg.double(), gt.double())
    ref.backward()
    rg, ag = render.cuda().requires_grad_(True), alpha.cuda().requires_grad_(True)
    out = blend_l1_ssim_loss(rg, ag, bg.cuda(), gt.cuda())
    (out * 1.7).backward()
    assert abs(float(out.detach()) - float(ref.detach())) < 1e-5 * max(1.0, abs(float(ref.detach())))
    want_r = ro.grad.clone() * 1.7
    if rs > 3:
        assert (rg.grad[..., 3:] == 0).all()
    assert grad_rel_err(rg.grad, want_r) < 1e-3
    assert grad_rel_err(ag.grad, ao.grad * 1.7) < 1e-3


@pytest.mark.gpu
def test_fused_loss_with_a_mask_matches_the_reference_lines(built_lib):
    """get_loss_dict multiplies gt and pred by batch["mask"] before L1 and SSIM (freegaussian_model.py:957-963)."""
    from freegaussian_b200.losses import blend_l1_ssim_loss
    H, W = 90, 131
    render, alpha, bg, gt = _inputs(H, W, 77, 4)
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(H, W, 1, generator=g) > 0.35)
    ro, ao = render.clone().double().requires_grad_(True), alpha.clone().double().requires_grad_(True)
    ref = OL.blend_l1_ssim_loss(ro[0], ao[0], bg.double(), gt.double(), mask=mask.double())
    ref.backward()
    rg, ag = render.cuda().requires_grad_(True), alpha.cuda().requires_grad_(True)
    out = blend_l1_ssim_loss(rg, ag, bg.cuda(), gt.cuda(), mask=mask.cuda())
    out.backward()
    assert abs(float(out.detach()) - float(ref.detach())) < 1e-5
    assert grad_rel_err(rg.grad, ro.grad) < 1e-3 and grad_rel_err(ag.grad, ao.grad) < 1e-3
    unmasked = blend_l1_ssim_loss(render.cuda(), alpha.cuda(), bg.cuda(), gt.cuda())
    assert abs(float(unmasked) - float(out.detach())) > 1e-3


@pytest.mark.gpu
def test_depth_fixup_matches_the_reference_lines(built_lib):
    """depth = where(alpha > 0, ED, ED.detach().max()) (freegaussian_model.py:884-886), forward bit-exact and backward."""
    from freegaussian_b200.losses import depth_fixup
    from freegaussian_b200.rendering import rasterization
    from util import small_scene
    W, H = 160, 96
    sc = small_scene(60, W, H, views=1, seed=6, scale_mul=0.1).to("cuda")  # sparse: ~40 % of the pixels have alpha == 0
    means = sc.means.clone().requires_grad_(True)
    render, alpha, _ = rasterization(means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, sc.Ks, W, H, packed=False,
                                     render_mode="RGB+ED", sh_degree=3)
    assert 0.05 < float((alpha == 0).float().mean()) < 0.95
    got = depth_fixup(render, alpha)
    want = OL.depth_fixup(render.detach().cpu(), alpha.detach().cpu())
    assert got.shape == want.shape == (1, H, W, 1) and torch.equal(got.detach().cpu(), want)
    # backward: gradient passes where alpha > 0 only, the maximum is a constant
    r2 = render.detach().clone().requires_grad_(True)
    w = torch.randn(1, H, W, 1, device="cuda")
    (depth_fixup(r2, alpha.detach()) * w).sum().backward()
    r3 = render.detach().cpu().clone().requires_grad_(True)
    (OL.depth_fixup(r3, alpha.detach().cpu()) * w.cpu()).sum().backward()
    assert torch.equal(r2.grad.cpu(), r3.grad)
    # and it chains into the renderer
    (got * w).sum().backward()
    assert means.grad is not None and float(means.grad.abs().max()) > 0
