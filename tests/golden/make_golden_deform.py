#!/usr/bin/env python
"""Generate tests/golden/deform_*.npz by RUNNING THE REFERENCE'S OWN CODE (this container only).

    python tests/golden/make_golden_deform.py            # needs /root/reference

The deformation network and its helpers are plain torch, but their modules import nerfstudio / gsplat at
the top, which are not installable here.  So:

* ``freegaussian/utils.py`` is imported as is, with one stub: ``nerfstudio.utils.misc.torch_compile``
  (a decorator the functions used here do not carry).
* ``FreeGaussianDeformableModel`` (and ``FreeGaussianControllableModel``) are cut out of
  ``freegaussian/freegaussian_model.py`` with ``ast`` and executed unmodified against those helpers.
* the application of the outputs (``freegaussian_model.py:836-845``) is a method body that cannot be cut
  out; it is re-typed below from those lines using the reference's ``to_homogenous`` / ``from_homogenous``.

Weights come from ``oracle.deform.init_params`` (seeded numpy draw) and are loaded with ``load_state_dict``,
so the fixtures store inputs, outputs, the input gradients and, of every weight gradient, the sum, the norm and
either all of it (up to 4096 values: biases, heads) or a strided sample.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import deform as OD  # noqa: E402

REF = "/root/reference/freegaussian"
OUT = os.path.dirname(os.path.abspath(__file__))
GRAD_STRIDE = 97


def load_reference():
    stub = types.ModuleType("nerfstudio.utils.misc")
    stub.torch_compile = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    for name in ("nerfstudio", "nerfstudio.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nerfstudio.utils.misc"] = stub
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "utils.py"))
    utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(utils)
    src = open(os.path.join(REF, "freegaussian_model.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "F": F, "get_embedder": utils.get_embedder, "exp_se3": utils.exp_se3}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("FreeGaussianDeformableModel", "FreeGaussianControllableModel"):
            exec(compile(ast.Module([node], []), "freegaussian_model.py", "exec"), ns)
    return utils, ns["FreeGaussianDeformableModel"], ns["FreeGaussianControllableModel"]


def scene(n, seed):
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(n, 3, generator=g) - 0.5) * 6.0  # model.py:155 with random_scale 6
    scales_log = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005)
    quats = torch.randn(n, 4, generator=g)
    weights = [torch.randn(n, k, generator=g) for k in (3, 3, 4)]  # fixed linear loss over (means, scales, quats)
    return means, scales_log, quats, weights


def deform_fixture(utils, Deform, name, n, seed, is_blender, t_val, scale):
    params = OD.init_params(is_blender=is_blender, seed=seed, scale=scale)
    net = Deform(is_blender=is_blender)
    net.load_state_dict(params, strict=True)
    means, scales_log, quats, wts = scene(n, seed)
    means.requires_grad_(True), scales_log.requires_grad_(True), quats.requires_grad_(True)
    t = torch.tensor([[t_val]]).expand(n, -1)
    # freegaussian_model.py:836-845
    d_xyz, d_rotation, d_scaling = net(means.detach(), t)
    new_means = utils.from_homogenous(torch.bmm(d_xyz, utils.to_homogenous(means).unsqueeze(-1)).squeeze(-1))
    new_scales = torch.exp(scales_log) + d_scaling
    new_quats = quats / quats.norm(dim=-1, keepdim=True) + d_rotation
    loss = (new_means * wts[0]).sum() + (new_scales * wts[1]).sum() + (new_quats * wts[2]).sum()
    loss.backward()
    # weight gradients: a strided sample, the sum and the norm of each (the full set is 0.6 M floats per fixture)
    grads = sampled_grads(net)
    np.savez_compressed(
        os.path.join(OUT, f"deform_{name}.npz"), n=n, seed=seed, is_blender=is_blender, t=t_val, scale=scale,
        means=means.detach().numpy(), scales_log=scales_log.detach().numpy(), quats=quats.detach().numpy(),
        w_means=wts[0].numpy(), w_scales=wts[1].numpy(), w_quats=wts[2].numpy(),
        d_xyz=d_xyz.detach().numpy(), d_rotation=d_rotation.detach().numpy(), d_scaling=d_scaling.detach().numpy(),
        new_means=new_means.detach().numpy(), new_scales=new_scales.detach().numpy(), new_quats=new_quats.detach().numpy(),
        grad_means=means.grad.numpy(), grad_scales_log=scales_log.grad.numpy(), grad_quats=quats.grad.numpy(), **grads)


def sampled_grads(net):
    grads = {}
    for k, v in net.named_parameters():
        g = v.grad.double().flatten()
        grads["grad." + k] = (g if g.numel() <= 4096 else g[::GRAD_STRIDE]).float().numpy()
        grads["gsum." + k] = np.float64(g.sum())
        grads["gnorm." + k] = np.float64(g.norm())
    return grads


def control_fixture(Control, name, n, seed):
    """FreeGaussianControllableModel (stage 2): x keeps its gradient (freegaussian_control_model.py:122, 143)."""
    net = Control()
    net.load_state_dict(OD.init_control_params(seed=seed), strict=True)
    g = torch.Generator().manual_seed(seed)
    x = ((torch.rand(n, 3, generator=g) - 0.5) * 6.0).requires_grad_(True)
    value = torch.randn(n, 3, generator=g) * 0.1
    wts = [torch.randn(n, k, generator=g) for k in (3, 4, 3)]
    d_xyz, d_rot, d_scale = net(x, value)
    ((d_xyz * wts[0]).sum() + (d_rot * wts[1]).sum() + (d_scale * wts[2]).sum()).backward()
    np.savez_compressed(os.path.join(OUT, f"control_{name}.npz"), n=n, seed=seed, x=x.detach().numpy(), value=value.numpy(),
                        w_xyz=wts[0].numpy(), w_rot=wts[1].numpy(), w_scale=wts[2].numpy(),
                        d_xyz=d_xyz.detach().numpy(), d_rot=d_rot.detach().numpy(), d_scale=d_scale.detach().numpy(),
                        grad_x=x.grad.numpy(), **sampled_grads(net))


if __name__ == "__main__":
    utils, Deform, Control = load_reference()
    deform_fixture(utils, Deform, "blender_n300", 300, 3, True, 0.37, 1.0)
    deform_fixture(utils, Deform, "blender_n129_hot", 129, 4, True, 0.81, 2.0)   # larger weights: bigger motions
    deform_fixture(utils, Deform, "real_n200", 200, 5, False, 0.55, 1.0)
    control_fixture(Control, "n250", 250, 6)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith(("deform_", "control_"))))
