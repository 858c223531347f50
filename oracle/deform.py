"""Oracle for the deformation network that precedes every render call.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  A plain-torch CPU restatement of

* ``Embedder`` / ``get_embedder``                       ``freegaussian/utils.py:8-56``
* ``skew`` / ``exp_so3`` / ``exp_se3``                  ``freegaussian/utils.py:82-159``
* ``FreeGaussianDeformableModel.forward``               ``freegaussian/freegaussian_model.py:1054-1114``
* ``FreeGaussianControllableModel.forward`` (stage 2)   ``freegaussian/freegaussian_model.py:1117-1145``
* the application of its outputs in ``get_outputs``     ``freegaussian/freegaussian_model.py:832-845``

PARITY PINNED: the logic lives in the reference repository itself (no un-vendored dependency), and
``tests/golden/deform_*.npz`` / ``control_*.npz`` hold outputs and gradients produced by executing the reference's own
class and function bodies (``tests/golden/make_golden_deform.py`` extracts them from
``/root/reference`` with ``ast``; only ``nerfstudio``'s ``torch_compile`` decorator import is stubbed).
``tests/test_deform.py`` checks this restatement against those fixtures.

Parameters use the reference's ``state_dict`` names (``linear.{i}.weight``, ``timenet.0.weight``,
``branch_w.weight``, ...), so a checkpoint of the reference module is a valid ``params`` dict.
"""

from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import Tensor


def embed(x: Tensor, multires: int) -> Tensor:
    """utils.py:27-56 with include_input, log sampling, [sin, cos]: [x, sin(x f0), cos(x f0), sin(x f1), ...]."""
    freqs = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)
    out = [x]
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def skew(w: Tensor) -> Tensor:
    """utils.py:82-96."""
    z = torch.zeros_like(w[:, 0])
    return torch.stack([z, -w[:, 2], w[:, 1], w[:, 2], z, -w[:, 0], -w[:, 1], w[:, 0], z], -1).reshape(-1, 3, 3)


def exp_se3(S: Tensor, theta: Tensor) -> Tensor:
    """utils.py:137-159 (with exp_so3, :116-134, and rp_to_se3, :99-113).  S [N,6], theta [N,1] -> [N,4,4]."""
    w, v = S[:, :3], S[:, 3:]
    W = skew(w)
    eye = torch.eye(3, dtype=S.dtype).expand(W.shape[0], 3, 3)
    W2 = torch.bmm(W, W)
    th = theta.view(-1, 1, 1)
    R = eye + torch.sin(th) * W + (1.0 - torch.cos(th)) * W2
    p = torch.bmm(th * eye + (1.0 - torch.cos(th)) * W + (th - torch.sin(th)) * W2, v.unsqueeze(-1))
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=S.dtype).repeat(W.shape[0], 1, 1)
    return torch.cat([torch.cat([R, p], -1), bottom], 1)


def time_embedding(params: Dict[str, Tensor], t: Tensor, is_blender: bool) -> Tensor:
    """model.py:1060, 1069-1071, 1092-1094: embedded time, passed through ``timenet`` when is_blender."""
    t_emb = embed(t, 6 if is_blender else 10)
    if is_blender:
        h = torch.relu(t_emb @ params["timenet.0.weight"].T + params["timenet.0.bias"])
        t_emb = h @ params["timenet.2.weight"].T + params["timenet.2.bias"]
    return t_emb


def trunk(params: Dict[str, Tensor], x: Tensor, t: Tensor, is_blender: bool, D: int = 8, multires: int = 10) -> Tensor:
    """model.py:1091-1101: the D-layer ReLU MLP with the skip connection after layer D//2.  Returns h [N,W]."""
    t_emb = time_embedding(params, t, is_blender)
    x_emb = embed(x, multires)
    h = torch.cat([x_emb, t_emb], -1)
    for i in range(D):
        h = torch.relu(h @ params[f"linear.{i}.weight"].T + params[f"linear.{i}.bias"])
        if i == D // 2:
            h = torch.cat([x_emb, t_emb, h], -1)
    return h


def deform_forward(params: Dict[str, Tensor], x: Tensor, t: Tensor, is_blender: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """model.py:1091-1114.  x [N,3], t [N,1] -> (d_xyz [N,4,4], rotation [N,4], scaling [N,3])."""
    h = trunk(params, x, t, is_blender)
    w = h @ params["branch_w.weight"].T + params["branch_w.bias"]
    v = h @ params["branch_v.weight"].T + params["branch_v.bias"]
    theta = torch.norm(w, dim=-1, keepdim=True)
    w = w / theta + 1e-5
    v = v / theta + 1e-5
    d_xyz = exp_se3(torch.cat([w, v], -1), theta)
    scaling = h @ params["gaussian_scaling.weight"].T + params["gaussian_scaling.bias"]
    rotation = h @ params["gaussian_rotation.weight"].T + params["gaussian_rotation.bias"]
    return d_xyz, rotation, scaling


def deform_gaussians(params: Dict[str, Tensor], means: Tensor, scales_log: Tensor, quats: Tensor, t: Tensor,
                     is_blender: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """model.py:836-845: the network sees ``means.detach()``; the SE(3) transform is applied to ``means``.

    Returns the (means, scales, quats) handed to ``rasterization`` at model.py:847-850.
    """
    d_xyz, d_rotation, d_scaling = deform_forward(params, means.detach(), t, is_blender)
    hom = torch.cat([means, torch.ones_like(means[..., :1])], -1)  # utils.py:59-68
    mh = torch.bmm(d_xyz, hom.unsqueeze(-1)).squeeze(-1)
    new_means = mh[..., :3] / mh[..., -1:]  # utils.py:71-80
    new_scales = torch.exp(scales_log) + d_scaling
    new_quats = quats / quats.norm(dim=-1, keepdim=True) + d_rotation
    return new_means, new_scales, new_quats


def control_forward(params: Dict[str, Tensor], x: Tensor, value: Tensor, D: int = 8, multires: int = 10
                    ) -> Tuple[Tensor, Tensor, Tensor]:
    """FreeGaussianControllableModel.forward, model.py:1135-1145: (d_xyz [N,3], d_rot [N,4], d_scale [N,3])."""
    v_emb = embed(value, multires)
    x_emb = embed(x, multires)
    h = torch.cat([x_emb, v_emb], -1)
    for i in range(D):
        h = torch.relu(h @ params[f"linear.{i}.weight"].T + params[f"linear.{i}.bias"])
        if i == D // 2:
            h = torch.cat([x_emb, v_emb, h], -1)
    lin = lambda n: h @ params[n + ".weight"].T + params[n + ".bias"]  # noqa: E731
    return lin("d_xyz"), lin("d_rot"), lin("d_scale")


def init_control_params(seed: int = 0, D: int = 8, W: int = 256, multires: int = 10, scale: float = 1.0) -> Dict[str, Tensor]:
    """Seeded parameters of the stage-2 control network (same recipe as ``init_params``)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    in_ch = 2 * (3 + 3 * 2 * multires)
    shapes = {}
    for i in range(D):
        shapes[f"linear.{i}"] = (W, in_ch if i == 0 else (W + in_ch if i == D // 2 + 1 else W))
    for name, o in (("d_xyz", 3), ("d_scale", 3), ("d_rot", 4)):
        shapes[name] = (o, W)
    out = {}
    for name, (o, i) in shapes.items():
        b = scale / np.sqrt(i)
        out[name + ".weight"] = torch.from_numpy(rng.uniform(-b, b, size=(o, i)).astype(np.float32))
        out[name + ".bias"] = torch.from_numpy(rng.uniform(-b, b, size=(o,)).astype(np.float32))
    return out


def init_params(is_blender: bool = True, seed: int = 0, D: int = 8, W: int = 256, multires: int = 10,
                scale: float = 1.0) -> Dict[str, Tensor]:
    """Seeded parameters with nn.Linear's shapes and default bound (U(-1/sqrt(in), 1/sqrt(in))), drawn with numpy's
    PCG64 so that the same values can be rebuilt anywhere (the fixtures do not have to store 0.6 M weights)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    x_ch = 3 + 3 * 2 * multires
    t_ch = 30 if is_blender else 1 + 2 * multires
    shapes = {}
    if is_blender:
        shapes["timenet.0"] = (256, 1 + 2 * 6)
        shapes["timenet.2"] = (30, 256)
    for i in range(D):
        fan_in = x_ch + t_ch if i == 0 else (W + x_ch + t_ch if i == D // 2 + 1 else W)
        shapes[f"linear.{i}"] = (W, fan_in)
    for name, o in (("branch_w", 3), ("branch_v", 3), ("gaussian_rotation", 4), ("gaussian_scaling", 3)):
        shapes[name] = (o, W)
    out = {}
    for name, (o, i) in shapes.items():
        b = scale / np.sqrt(i)
        out[name + ".weight"] = torch.from_numpy(rng.uniform(-b, b, size=(o, i)).astype(np.float32))
        out[name + ".bias"] = torch.from_numpy(rng.uniform(-b, b, size=(o,)).astype(np.float32))
    return out
