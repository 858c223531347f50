"""The tiny pure helpers the reference imports next to ``rasterization``.

* ``num_sh_bases``  -- ``gsplat.cuda_legacy._wrapper.num_sh_bases``
  (``freegaussian_model.py:21``, used ``:165``).
* ``quat_to_rotmat`` -- ``gsplat.cuda_legacy._torch_impl.quat_to_rotmat``
  (``freegaussian_model.py:15``, used ``:535`` when splitting Gaussians).
* ``get_viewmat``   -- the camera convention of ``freegaussian/utils.py:162-179``
  (SURVEY.md section 8(a) row a7): OpenGL camera-to-world -> OpenCV world-to-camera.

They are O(N) elementwise torch expressions evaluated on whatever device the inputs
live on; none of them is on the kernel path.
"""

from __future__ import annotations

import torch
from torch import Tensor


def num_sh_bases(degree: int) -> int:
    assert 0 <= degree <= 4, "SH degree must be in 0..4"
    return (degree + 1) ** 2


def quat_to_rotmat(quat: Tensor) -> Tensor:
    """[...,4] (w,x,y,z), normalised inside -> [...,3,3]."""
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(torch.nn.functional.normalize(quat, dim=-1), dim=-1)
    mat = torch.stack(
        [
            1 - 2 * (y**2 + z**2), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x**2 + z**2), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x**2 + y**2),
        ],
        dim=-1,
    )
    return mat.reshape(quat.shape[:-1] + (3, 3))


def get_viewmat(camera_to_world: Tensor) -> Tensor:
    """[C,3,4] or [C,4,4] OpenGL c2w -> [C,4,4] world-to-camera with y/z flipped."""
    R = camera_to_world[:, :3, :3]
    T = camera_to_world[:, :3, 3:4]
    R = R * torch.tensor([[[1.0, -1.0, -1.0]]], device=R.device, dtype=R.dtype)
    R_inv = R.transpose(1, 2)
    T_inv = -torch.bmm(R_inv, T)
    viewmat = torch.zeros(R.shape[0], 4, 4, device=R.device, dtype=R.dtype)
    viewmat[:, 3, 3] = 1.0
    viewmat[:, :3, :3] = R_inv
    viewmat[:, :3, 3:4] = T_inv
    return viewmat
