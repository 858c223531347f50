"""Fused optimizer step for the Gaussian parameter groups.

The reference gives every group its own ``torch.optim.Adam`` (``freegaussian_config.py:48-75``:
``AdamOptimizerConfig(lr=..., eps=1e-15)``, default betas, no weight decay; only ``means`` has a
scheduler).  :class:`GaussianAdam` steps all of them with ONE kernel launch (``fg_adam_step``),
follows ``torch.optim.Adam``'s arithmetic operation by operation, and exposes ``exp_avg`` /
``exp_avg_sq`` per group so the refinement (``densify.refine``) can do the reference's state
surgery (``freegaussian_model.py:313-367``).

A group may hold two reference groups in one tensor: the ``[N,16,3]`` SH tensor the renderer
consumes is ``cat(features_dc[:,None], features_rest)`` (``freegaussian_model.py:801``); with
``split=3`` its first three columns step with ``lr`` (features_dc) and the rest with ``lr_rest``
(features_rest), so the per-step concatenation and its backward split are not needed.

Multi-GPU: ``step(shard=(rank, world))`` updates only this rank's contiguous slice of every group
(the flat gradient arena after a reduce-scatter, SURVEY 8(e) "better variant"); the caller
all-gathers the parameters afterwards.  The kernel side is tested (``tests/test_optim.py::test_sharded_step_equals_full_step``);
the reduce-scatter / all-gather plumbing around it is not built yet ("Not widened yet").
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check

# freegaussian_config.py:48-75 (Gaussian groups only)
REFERENCE_LRS = {
    "means": 1.6e-4 * 5,
    "features_dc": 0.0025,
    "features_rest": 0.0025 / 20,
    "opacities": 0.05,
    "scales": 0.001 * 5,
    "quats": 0.001,
}
MEANS_LR_FINAL = 1.6e-6 * 5  # freegaussian_config.py:52-55
MEANS_LR_MAX_STEPS = 30000


def exponential_decay_lr(step: int, lr_init: float, lr_final: float, max_steps: int) -> float:
    """nerfstudio ``ExponentialDecayScheduler`` without warm-up [upstream, un-vendored; nerfstudio>=1.1.3,
    ``pyproject.toml:9``]: log-linear interpolation from lr_init to lr_final over max_steps."""
    t = min(max(step / max_steps, 0.0), 1.0)
    return math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


@dataclass
class Group:
    param: Tensor
    lr: float
    lr_rest: float = 0.0
    split: int = 0  # columns [0,split) of each row use lr, the rest lr_rest; 0 = one rate
    exp_avg: Optional[Tensor] = None
    exp_avg_sq: Optional[Tensor] = None

    @property
    def row_len(self) -> int:
        p = self.param
        return int(p.numel() // p.shape[0]) if (self.split and p.shape[0]) else 0


class GaussianAdam:
    """``torch.optim.Adam``-equivalent over named groups, one launch per step."""

    def __init__(self, groups: Dict[str, Group], betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-15,
                 t: int = 0):
        assert len(groups) <= _lib.ADAM_MAX_SEGMENTS, "too many groups for one launch"
        self.groups = groups
        self.betas = betas
        self.eps = eps
        self.t = int(t)  # steps taken so far (bias correction); survives rebind()
        for name, g in groups.items():
            if not g.param.is_cuda:
                raise RuntimeError(f"GaussianAdam: group `{name}` is not a CUDA tensor (no CPU path)")
            assert g.param.dtype == torch.float32 and g.param.is_contiguous()
            if g.exp_avg is None:  # torch creates the state lazily with zeros
                g.exp_avg = torch.zeros_like(g.param)
                g.exp_avg_sq = torch.zeros_like(g.param)

    @classmethod
    def for_reference_groups(cls, means, sh, opacities, scales, quats) -> "GaussianAdam":
        """The six Gaussian groups with the reference's learning rates; ``sh`` is the [N,16,3] tensor."""
        R = REFERENCE_LRS
        return cls({
            "means": Group(means, R["means"]),
            "sh": Group(sh, R["features_dc"], R["features_rest"], split=3),
            "opacities": Group(opacities, R["opacities"]),
            "scales": Group(scales, R["scales"]),
            "quats": Group(quats, R["quats"]),
        })

    def rebind(self, params: Dict[str, Tensor], state: Optional[Dict[str, Tuple[Tensor, Tensor]]] = None) -> None:
        """Swap in the tensors a refinement returned (``densify.refine(...).params`` / ``.state``: rows were split,
        duplicated or culled, so every tensor is a new allocation) WITHOUT restarting the bias correction: the
        reference keeps ``step`` inside the surviving ``param_state`` (``freegaussian_model.py:313-367``,
        ``dup_in_optim`` / ``remove_from_optim``), so ``t`` carries on.  Groups missing from ``state`` get zero moments."""
        for name, g in self.groups.items():
            if name not in params:
                continue
            p = params[name]
            assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous(), name
            g.param = p
            if state is not None and name in state:
                m, v = state[name]
                assert m.shape == p.shape and v.shape == p.shape, name
                g.exp_avg, g.exp_avg_sq = m, v
            else:
                g.exp_avg, g.exp_avg_sq = torch.zeros_like(p), torch.zeros_like(p)

    def set_lr(self, name: str, lr: float, lr_rest: Optional[float] = None) -> None:
        self.groups[name].lr = lr
        if lr_rest is not None:
            self.groups[name].lr_rest = lr_rest

    @torch.no_grad()
    def step(self, grads: Optional[Dict[str, Tensor]] = None, shard: Optional[Tuple[int, int]] = None) -> None:
        """``grads[name]`` defaults to ``param.grad``.  Groups without a gradient are skipped (as torch does).
        One step counter serves every group (the reference steps all of its Gaussian optimizers every iteration, so
        their per-parameter ``step`` values coincide)."""
        dev = next(iter(self.groups.values())).param.device
        if dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return self.step(grads, shard)
        self.t += 1
        segs = (_lib.AdamSegment * _lib.ADAM_MAX_SEGMENTS)()
        k = 0
        keep = []
        for name, g in self.groups.items():
            grad = grads.get(name) if grads is not None else g.param.grad
            if grad is None or g.param.numel() == 0:
                continue
            assert grad.shape == g.param.shape and grad.dtype == torch.float32 and grad.is_cuda
            grad = grad.contiguous()
            keep.append(grad)
            n = g.param.numel()
            lo, hi = 0, n
            if shard is not None:
                lo, hi = shard_range(n, *shard)
            s = segs[k]
            s.param = g.param.data_ptr() + 4 * lo
            s.grad = grad.data_ptr() + 4 * lo
            s.exp_avg = g.exp_avg.data_ptr() + 4 * lo
            s.exp_avg_sq = g.exp_avg_sq.data_ptr() + 4 * lo
            s.n, s.first = hi - lo, lo
            s.row_len, s.split = g.row_len, g.split
            s.lr, s.lr_rest = g.lr, g.lr_rest
            k += 1
        check(_lib.lib().fg_adam_step(k, segs, self.t, self.betas[0], self.betas[1], self.eps,
                                      torch.cuda.current_stream().cuda_stream))

    def state(self) -> Dict[str, Tuple[Tensor, Tensor]]:
        return {n: (g.exp_avg, g.exp_avg_sq) for n, g in self.groups.items()}


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of n elements owned by `rank` (multiples of 4 floats so the slices stay 16-byte
    aligned; the last rank takes the remainder)."""
    per = (n // world) // 4 * 4
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi
