"""Seeded synthetic scenes of the BASELINE.json shapes.

Pure torch, device-agnostic, no dataset needed.  Two recipes:

* ``init_like``    -- exactly the reference's ``random_init`` recipe
  (``freegaussian_model.py:155-186``): uniform cube, scale = mean 3-NN distance,
  random unit quats (``utils.py:214-229``), opacity 0.1, DC colour ~ U(0,1).
* ``trained_like`` -- same cube, 80 % of the Gaussians on 6 planes + 20 % volume,
  ``opacity = sigmoid(N(0, 2^2))``, per-axis anisotropy ``exp(N(0, 0.5^2))``,
  ``features_rest ~ N(0, 0.05^2)``: realistic early termination.

Cameras sit on a ring looking at the origin in the OpenGL camera-to-world convention
and go through :func:`freegaussian_b200.compat.get_viewmat` like
``freegaussian_model.py:810``; intrinsics follow ``freegaussian_dataparser.py:1219-1224``.
The frame pair t/t+1 moves an "articulated part" (a sub-box with ~10 % of the
Gaussians) by a rigid screw motion.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional

import torch
from torch import Tensor

from .compat import get_viewmat

CAMERA_ANGLE_X = 0.6911  # rad; D-NeRF/LiveScene style transforms.json


@dataclass
class Scene:
    means: Tensor  # [N,3]
    quats: Tensor  # [N,4] (w,x,y,z), unit
    scales: Tensor  # [N,3] linear (already exp'ed, as at freegaussian_model.py:844)
    opacities: Tensor  # [N] in (0,1) (already sigmoid'ed, freegaussian_model.py:851)
    sh: Tensor  # [N,16,3] (features_dc ++ features_rest, freegaussian_model.py:801)
    viewmats: Tensor  # [C,4,4] world->camera
    Ks: Tensor  # [C,3,3]
    width: int
    height: int
    means_next: Tensor  # [N,3] frame t+1
    quats_next: Tensor  # [N,4] frame t+1

    def to(self, device) -> "Scene":
        kw = {k: (v.to(device) if isinstance(v, Tensor) else v) for k, v in self.__dict__.items()}
        return Scene(**kw)


def random_quats(n: int, gen: torch.Generator) -> Tensor:
    """Uniform random unit quaternions, same recipe as ``utils.py:214-229``."""
    u, v, w = (torch.rand(n, generator=gen) for _ in range(3))
    return torch.stack(
        [
            torch.sqrt(1 - u) * torch.sin(2 * math.pi * v),
            torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
            torch.sqrt(u) * torch.sin(2 * math.pi * w),
            torch.sqrt(u) * torch.cos(2 * math.pi * w),
        ],
        dim=-1,
    )


def ring_cameras(n_views: int, extent: float, width: int, height: int, phase: float = 0.0):
    """C views on a ring of radius 0.75*extent at height 0.3*extent looking at the origin."""
    r, h = 1.5 * extent / 2, 0.3 * extent
    c2w = torch.zeros(n_views, 3, 4)
    for i in range(n_views):
        th = phase + 2 * math.pi * i / max(n_views, 1) + 0.35
        eye = torch.tensor([r * math.cos(th), r * math.sin(th), h])
        fwd = -eye / eye.norm()  # camera looks along -z_cam (OpenGL)
        up = torch.tensor([0.0, 0.0, 1.0])
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        up2 = torch.linalg.cross(right, fwd)
        c2w[i, :, 0], c2w[i, :, 1], c2w[i, :, 2], c2w[i, :, 3] = right, up2, -fwd, eye
    viewmats = get_viewmat(c2w)
    f = 0.5 * width / math.tan(0.5 * CAMERA_ANGLE_X)
    K = torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]])
    return viewmats, K[None].repeat(n_views, 1, 1)


def quat_mul(a: Tensor, b: Tensor) -> Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        -1,
    )


def make_scene(
    n: int,
    width: int,
    height: int,
    n_views: int = 1,
    recipe: str = "trained_like",
    extent: float = 6.0,
    seed: int = 0,
    knn3: Optional[Callable[[Tensor], Tensor]] = None,
    view_phase: float = 0.0,
) -> Scene:
    """Build a seeded scene.  ``knn3(means) -> [N,3]`` distances to the 3 nearest
    neighbours (self excluded); it is the caller's KNN (GPU kernel in bench.py, the
    sklearn oracle in CPU tests) so this module has no CPU compute fallback."""
    assert recipe in ("init_like", "trained_like")
    gen = torch.Generator().manual_seed(seed)
    if recipe == "init_like":
        means = (torch.rand(n, 3, generator=gen) - 0.5) * extent
    else:
        n_surf = int(0.8 * n)
        pts = (torch.rand(n, 3, generator=gen) - 0.5) * extent
        plane = torch.randint(0, 6, (n_surf,), generator=gen)
        axis = plane % 3
        offs = torch.tensor([-0.25, 0.25])[(plane // 3)] * extent
        jitter = torch.randn(n_surf, generator=gen) * 0.002 * extent
        pts[torch.arange(n_surf), axis] = offs + jitter
        means = pts[torch.randperm(n, generator=gen)]
    assert knn3 is not None, "pass knn3= (GPU k-NN in bench.py, sklearn oracle in tests)"
    d3 = knn3(means).to(torch.float32).cpu()
    base = d3.mean(-1, keepdim=True).clamp_min(1e-6).repeat(1, 3)
    quats = random_quats(n, gen)
    sh = torch.zeros(n, 16, 3)
    sh[:, 0] = torch.rand(n, 3, generator=gen)
    if recipe == "init_like":
        scales = base
        opac = torch.full((n,), 0.1)
    else:
        scales = base * torch.exp(torch.randn(n, 3, generator=gen) * 0.5)
        opac = torch.sigmoid(torch.randn(n, generator=gen) * 2.0)
        sh[:, 0] = (sh[:, 0] - 0.5) / 0.28209479177387814  # RGB2SH, utils.py:232-237
        sh[:, 1:] = torch.randn(n, 15, 3, generator=gen) * 0.05
    viewmats, Ks = ring_cameras(n_views, extent, width, height, view_phase)

    # articulated part: sub-box with ~10 % of the volume's Gaussians, screw motion
    half = 0.5 * extent * (0.1 ** (1 / 3))
    centre = torch.tensor([0.1, -0.05, 0.0]) * extent
    part = ((means - centre).abs() < half).all(-1)
    ang = math.radians(2.0)
    axis = torch.tensor([0.0, 0.0, 1.0])
    Rz = torch.tensor(
        [[math.cos(ang), -math.sin(ang), 0.0], [math.sin(ang), math.cos(ang), 0.0], [0.0, 0.0, 1.0]]
    )
    moved = (means - centre) @ Rz.T + centre + 0.01 * extent * axis
    means_next = torch.where(part[:, None], moved, means)
    dq = torch.tensor([math.cos(ang / 2), 0.0, 0.0, math.sin(ang / 2)])
    quats_next = torch.where(part[:, None], quat_mul(dq.expand(n, 4), quats), quats)
    return Scene(means, quats, scales, opac, sh, viewmats, Ks, width, height, means_next, quats_next)
