"""Oracle for the optimizer step of the Gaussian parameter groups (``freegaussian_config.py:48-75``).

TEST INFRASTRUCTURE ONLY.  The reference steps each group with ``torch.optim.Adam(lr, eps=1e-15)``
(nerfstudio ``AdamOptimizerConfig`` [upstream, un-vendored]; torch is importable here and on the GPU
box).  PARITY PINNED: ``adam_reference`` below IS ``torch.optim.Adam`` run on CPU tensors;
``adam_step_restated`` spells the same update out and is checked against it in
``tests/test_optim.py``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
from torch import Tensor


def adam_reference(params: Dict[str, Tensor], grads: List[Dict[str, Tensor]], lrs: Dict[str, float],
                   eps: float = 1e-15) -> Tuple[Dict[str, Tensor], Dict[str, Tuple[Tensor, Tensor]]]:
    """One torch.optim.Adam per group (as nerfstudio's Optimizers builds them), stepped len(grads) times."""
    ps = {k: torch.nn.Parameter(v.detach().clone()) for k, v in params.items()}
    opts = {k: torch.optim.Adam([ps[k]], lr=lrs[k], eps=eps) for k in ps}
    for g in grads:
        for k in ps:
            ps[k].grad = g[k].clone()
            opts[k].step()
    state = {k: (opts[k].state[ps[k]]["exp_avg"], opts[k].state[ps[k]]["exp_avg_sq"]) for k in ps}
    return {k: v.detach() for k, v in ps.items()}, state


def adam_step_restated(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float = 0.9,
                       beta2: float = 0.999, eps: float = 1e-15) -> None:
    """torch/optim/adam.py `_single_tensor_adam` (no amsgrad, no weight decay, not maximize), in place."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1
    denom = (v.sqrt() / math.sqrt(bias_correction2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


def exponential_decay_lr(step: int, lr_init: float, lr_final: float, max_steps: int) -> float:
    """nerfstudio ExponentialDecayScheduler without warm-up [upstream]: log-linear interpolation."""
    t = min(max(step / max_steps, 0.0), 1.0)
    return math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
