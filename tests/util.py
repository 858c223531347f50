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
