"""Oracle for the refinement step (``freegaussian_model.py:404-571`` with the optimizer-state
surgery of ``:313-367``).

TEST INFRASTRUCTURE ONLY.  PARITY PINNED: the logic is plain torch indexing that lives in the
reference repository itself; it is restated here statement by statement on CPU tensors, with the
model's attributes replaced by a ``params`` dict and the optimizer states by ``state[name] =
[exp_avg, exp_avg_sq]``.  ``quat_to_rotmat`` is gsplat's legacy helper [upstream] as restated in
``freegaussian_b200/compat.py`` (normalise, then the standard wxyz formula).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor


def quat_to_rotmat(quat: Tensor) -> Tensor:
    quat = torch.nn.functional.normalize(quat, dim=-1)
    w, x, y, z = torch.unbind(quat, dim=-1)
    m = torch.stack([
        1 - 2 * (y ** 2 + z ** 2), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x ** 2 + z ** 2), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x ** 2 + y ** 2),
    ], dim=-1)
    return m.reshape(quat.shape[:-1] + (3, 3))


class RefineOracle:
    """Holds what `refinement_after` touches on the model: params, Adam moments, statistics, config."""

    def __init__(self, params: Dict[str, Tensor], state: Dict[str, List[Tensor]], cfg, step: int,
                 num_train_data: int, last_size: Tuple[int, int], xys_grad_norm: Optional[Tensor],
                 vis_counts: Optional[Tensor], max_2Dsize: Optional[Tensor], samples: Optional[Tensor] = None):
        self.p = {k: v.clone() for k, v in params.items()}
        self.state = {k: [m.clone(), v.clone()] for k, (m, v) in state.items()}
        self.cfg, self.step, self.num_train_data, self.last_size = cfg, step, num_train_data, last_size
        self.xys_grad_norm, self.vis_counts, self.max_2Dsize = xys_grad_norm, vis_counts, max_2Dsize
        self.samples = samples
        self.opacity_reset = False

    @property
    def num_points(self) -> int:
        return self.p["means"].shape[0]

    # :313-338
    def remove_from_all_optim(self, deleted_mask: Tensor) -> None:
        for k in self.state:
            self.state[k] = [self.state[k][0][~deleted_mask], self.state[k][1][~deleted_mask]]

    # :340-367
    def dup_in_all_optim(self, idcs: Tensor, n: int) -> None:
        for k in self.state:
            for j in range(2):
                s = self.state[k][j]
                rep = (n,) + tuple(1 for _ in range(s.dim() - 1))
                self.state[k][j] = torch.cat([s, torch.zeros_like(s[idcs]).repeat(*rep)], dim=0)

    # :513-556
    def split_gaussians(self, split_mask: Tensor, samps: int) -> Dict[str, Tensor]:
        n_splits = int(split_mask.sum().item())
        centered = self.samples if self.samples is not None else torch.randn((samps * n_splits, 3))  # :519
        assert centered.shape == (samps * n_splits, 3)
        scaled = torch.exp(self.p["scales"][split_mask].repeat(samps, 1)) * centered  # :520-522
        quats = self.p["quats"][split_mask] / self.p["quats"][split_mask].norm(dim=-1, keepdim=True)  # :523
        rots = quat_to_rotmat(quats.repeat(samps, 1))  # :524
        rotated = torch.bmm(rots, scaled[..., None]).squeeze(-1)  # :525
        out = {"means": rotated + self.p["means"][split_mask].repeat(samps, 1)}  # :526
        size_fac = 1.6
        out["scales"] = torch.log(torch.exp(self.p["scales"][split_mask]) / size_fac).repeat(samps, 1)  # :535
        self.p["scales"][split_mask] = torch.log(torch.exp(self.p["scales"][split_mask]) / size_fac)  # :536
        for name, param in self.p.items():  # :528-533, :538, :547-549
            if name not in out:
                rep = (samps,) + tuple(1 for _ in range(param.dim() - 1))
                out[name] = param[split_mask].repeat(*rep)
        return out

    # :558-567
    def dup_gaussians(self, dup_mask: Tensor) -> Dict[str, Tensor]:
        return {name: param[dup_mask] for name, param in self.p.items()}

    # :493-511
    def cull_gaussians(self, extra_cull_mask: Optional[Tensor] = None) -> Tensor:
        c = self.cfg
        culls = (torch.sigmoid(self.p["opacities"]) < c.cull_alpha_thresh).squeeze(-1)
        if extra_cull_mask is not None:
            culls = culls | extra_cull_mask
        if self.step > c.refine_every * c.reset_alpha_every:
            toobigs = (torch.exp(self.p["scales"]).max(dim=-1).values > c.cull_scale_thresh)
            if self.step < c.stop_screen_size_at:
                if self.max_2Dsize is not None:
                    toobigs = toobigs | (self.max_2Dsize > c.cull_screen_size)
            culls = culls | toobigs
        for name, param in self.p.items():
            self.p[name] = param[~culls]
        return culls

    # :404-491
    def refinement_after(self) -> None:
        c = self.cfg
        if self.step < c.refine_start:
            return
        reset_interval = c.reset_alpha_every * c.refine_every
        do_densification = (self.step < c.stop_split_at
                            and self.step % reset_interval > self.num_train_data + c.refine_every)
        if do_densification:
            avg_grad_norm = (self.xys_grad_norm / self.vis_counts) * 0.5 * max(self.last_size[0], self.last_size[1])
            high_grads = avg_grad_norm > c.densify_grad_thresh
            splits = self.p["scales"].exp().max(dim=-1).values > c.densify_size_thresh
            splits &= high_grads
            if self.step < c.stop_screen_size_at:
                splits |= self.max_2Dsize > c.split_screen_size
            nsamps = c.n_split_samples
            split_params = self.split_gaussians(splits, nsamps)
            dups = self.p["scales"].exp().max(dim=-1).values <= c.densify_size_thresh  # after :536 rescaled the parents
            dups &= high_grads
            dup_params = self.dup_gaussians(dups)
            for name, param in self.p.items():
                self.p[name] = torch.cat([param, split_params[name], dup_params[name]], dim=0)
            self.max_2Dsize = torch.cat([self.max_2Dsize, torch.zeros_like(split_params["scales"][:, 0]),
                                         torch.zeros_like(dup_params["scales"][:, 0])], dim=0)
            self.dup_in_all_optim(torch.where(splits)[0], nsamps)
            self.dup_in_all_optim(torch.where(dups)[0], 1)
            splits_mask = torch.cat((splits, torch.zeros(nsamps * int(splits.sum()) + int(dups.sum()), dtype=torch.bool)))
            deleted_mask = self.cull_gaussians(splits_mask)
            self.n_split, self.n_dup = int(splits.sum()), int(dups.sum())
        elif self.step >= c.stop_split_at and c.continue_cull_post_densification:
            deleted_mask = self.cull_gaussians()
        else:
            deleted_mask = None
        if deleted_mask is not None:
            self.remove_from_all_optim(deleted_mask)
        if self.step < c.stop_split_at and self.step % reset_interval == c.refine_every:
            reset_value = c.cull_alpha_thresh * 2.0
            self.p["opacities"] = torch.clamp(self.p["opacities"], max=torch.logit(torch.tensor(reset_value)).item())
            if "opacities" in self.state:
                self.state["opacities"] = [torch.zeros_like(s) for s in self.state["opacities"]]
            self.opacity_reset = True
        self.xys_grad_norm = None
        self.vis_counts = None
        self.max_2Dsize = None
