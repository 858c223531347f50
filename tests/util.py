"""Shared helpers for the parity tests (oracle side runs on CPU)."""
from __future__ import annotations

import torch

from freegaussian_b200.scenes import make_scene
from oracle import knn as oknn


def knn3_cpu(means: torch.Tensor) -> torch.Tensor:
    return torch.from_numpy(oknn.reference_knn(means.numpy(), 3)[0])


def small_scene(n=1500, w=96, h=64, views=2, recipe="trained_like", seed=0, scale_mul=1.0):
    sc = make_scene(n, w, h, n_views=views, recipe=recipe, seed=seed, knn3=knn3_cpu)
    sc.scales = sc.scales * scale_mul
    return sc


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the scale of the reference tensor b."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = max(1.0, float(b.abs().max())) if b.numel() else 1.0
    return float((a - b).abs().max()) / denom if b.numel() else 0.0


def grad_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Gradient tolerance metric: max abs error over the reference's max magnitude."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = float(b.abs().max())
    if denom == 0:
        return float(a.abs().max())
    return float((a - b).abs().max()) / denom


def oracle_window(leaf, viewmat, K, W, H, win, absgrad=False):
    """Composite a tile-aligned window ``win = (x0, y0, cw, ch)`` of a W x H frame with the CPU oracle from the Gaussians in
    ``leaf`` (dict: means, quats, scales, opacities, sh, means_next -- the subset that can reach the window): projection,
    SH, tile lists and sort at the FULL frame's camera (a shifted principal point would change gsplat's frustum clamp of the
    projection Jacobian), then the oracle's compositing over the window's tiles only.  Differentiable w.r.t. ``leaf``.
    Returns (render [1,ch,cw,4] with "ED" normalisation, alpha, flow, absgrad sink [n,2] | None)."""
    from oracle import render as O
    x0, y0, cw, ch = win
    assert x0 % 16 == 0 and y0 % 16 == 0 and cw % 16 == 0 and ch % 16 == 0
    means, quats, scales, opac, sh, mnext = (leaf[k] for k in ("means", "quats", "scales", "opacities", "sh", "means_next"))
    vm = viewmat
    radii, means2d, depths, conics, _, _ = O.fully_fused_projection(means, quats, scales, vm, K, W, H, 0.3, 0.01, 1e10, 0.0)
    vis = radii > 0
    dirs = means[None] - torch.inverse(vm)[:, :3, 3][:, None]
    cols = torch.clamp_min(O.spherical_harmonics(3, dirs, sh[None], masks=vis) + 0.5, 0.0)
    uv_next, z_next = O.project_points(mnext, vm, K)
    flow2d = torch.where((vis & (z_next >= 0.01))[..., None], uv_next - means2d, torch.zeros(()))
    cols = torch.cat([cols, depths[..., None], flow2d], -1)  # rgb | depth | flow, as oracle.rasterization builds them
    tile_w, tile_h = (W + 15) // 16, (H + 15) // 16
    _, isect_ids, flatten_ids = O.isect_tiles(means2d, radii, depths, 16, tile_w, tile_h)
    offs = O.isect_offset_encode(isect_ids, 1, tile_w, tile_h).reshape(-1).tolist() + [flatten_ids.numel()]
    win_ids, win_offs, total = [], [], 0
    for ty in range(y0 // 16, (y0 + ch) // 16):
        for tx in range(x0 // 16, (x0 + cw) // 16):
            t = ty * tile_w + tx
            win_offs.append(total)
            win_ids.append(flatten_ids[offs[t]:offs[t + 1]])
            total += win_ids[-1].numel()
    win_ids = torch.cat(win_ids)
    win_offs = torch.tensor(win_offs, dtype=torch.int32).view(1, ch // 16, cw // 16)
    shifted = means2d - torch.tensor([float(x0), float(y0)])
    sink = torch.zeros(means.shape[0], 2) if absgrad else None
    ro, ao, _ = O.rasterize_to_pixels(shifted, conics, cols, opac[None], cw, ch, 16, win_offs, win_ids, absgrad_sink=sink)
    render = torch.cat([ro[..., :3], ro[..., 3:4] / ao.clamp(min=1e-10)], -1)  # "ED" normalisation
    return render, ao, ro[..., 4:], sink, win_ids.numel()
