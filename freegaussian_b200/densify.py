"""Refinement of the Gaussian set -- split / duplicate / cull with the Adam-state surgery
 -- as three kernels instead of the reference's mask-index / ``torch.cat`` chain.

``refine(...)`` mirrors ``FreeGaussianModel.refinement_after`` (``freegaussian_model.py:404-491``):
same schedule tests, same masks, same output row order (kept originals, the children of split
Gaussians sample-major, duplicates), zeros in the Adam moments of new rows (``:341-367``), moments
of removed rows dropped (``:313-338``), opacity reset (``:470-483``).  The random draw of
``split_gaussians`` (``:519``) stays a ``torch.randn`` call of the same shape, so a run seeded like
the reference consumes the same numbers.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr


@dataclass
class RefineSchedule:
    """Defaults of ``FreeGaussianModelConfig`` (``freegaussian_model.py:58-99``)."""
    refine_start: int = 500
    refine_every: int = 100
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    continue_cull_post_densification: bool = True
    reset_alpha_every: int = 30
    densify_grad_thresh: float = 0.0008
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    cull_screen_size: float = 0.15
    split_screen_size: float = 0.05
    stop_screen_size_at: int = 4000
    stop_split_at: int = 15000


@dataclass
class RefineResult:
    params: Dict[str, Tensor]
    state: Dict[str, Tuple[Tensor, Tensor]]
    n_before: int
    n_after: int
    n_split: int
    n_dup_kept: int  # duplicates that survive the cull that follows (the reference logs the pre-cull count)
    n_kept: int
    src: Optional[Tensor]  # [n_after] int32 parent row of every output row (None if nothing changed)
    opacity_reset: bool


def plan_and_apply(params: Dict[str, Tensor], state: Dict[str, Tuple[Tensor, Tensor]], grad_norm: Optional[Tensor],
                   vis_count: Optional[Tensor], max_size: Optional[Tensor], *, densify: bool, use_screen: bool,
                   cull_big: bool, max_dim: float, cfg: RefineSchedule, samples: Optional[Tensor] = None,
                   generator: Optional[torch.Generator] = None) -> RefineResult:
    """The kernel pipeline: fg_refine_plan -> (one host read of four counts) -> fg_refine_map ->
    fg_refine_gather (every parameter and Adam moment in one launch) -> fg_refine_children."""
    L = _lib.lib()
    means, scales, quats, opac = params["means"], params["scales"], params["quats"], params["opacities"]
    for name, t in params.items():
        if not t.is_cuda:
            raise RuntimeError(f"refine: `{name}` is not a CUDA tensor (no CPU path)")
        assert t.dtype == torch.float32 and t.is_contiguous() and t.shape[0] == means.shape[0], name
    dev, N = means.device, int(means.shape[0])
    st = torch.cuda.current_stream().cuda_stream
    c = _lib.RefineConfig(cfg.densify_grad_thresh, cfg.densify_size_thresh, cfg.split_screen_size,
                          cfg.cull_alpha_thresh, cfg.cull_scale_thresh, cfg.cull_screen_size, float(max_dim),
                          cfg.n_split_samples, int(use_screen), int(cull_big), int(densify))
    flat = lambda t: None if t is None else t.reshape(-1).contiguous()
    plan = torch.empty(4 * N + 1, dtype=torch.int32, device=dev)
    counts = torch.empty(4, dtype=torch.int64, device=dev)
    ws = torch.empty(int(L.fg_refine_workspace_bytes(N)), dtype=torch.uint8, device=dev)
    check(L.fg_refine_plan(N, ptr(scales), ptr(flat(opac)), ptr(flat(grad_norm)), ptr(flat(vis_count)),
                           ptr(flat(max_size)), c, ptr(plan), ptr(counts), ptr(ws), ws.numel(), st))
    n_keep, n_kc, n_kd, n_split = (int(v) for v in counts.tolist())  # the one host sync of a refinement
    samps = cfg.n_split_samples
    n_children = samps * n_kc
    n_out = n_keep + n_children + n_kd
    if n_out == N and n_keep == N:
        return RefineResult(params, state, N, N, 0, 0, N, None, False)
    src = torch.empty(n_out, dtype=torch.int32, device=dev)
    sample_row = torch.empty(n_out, dtype=torch.int32, device=dev)
    check(L.fg_refine_map(N, ptr(plan), ptr(counts), samps, n_out, ptr(src), ptr(sample_row), st))
    arrays = (_lib.RefineArray * _lib.REFINE_MAX_ARRAYS)()
    new_params: Dict[str, Tensor] = {}
    new_state: Dict[str, Tuple[Tensor, Tensor]] = {}
    k = 0

    def add(t: Tensor, zero_new: bool) -> Tensor:
        nonlocal k
        assert k < _lib.REFINE_MAX_ARRAYS
        out = torch.empty((n_out,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
        row = t.numel() // max(N, 1)
        if n_out and row:
            arrays[k].in_, arrays[k].out = t.data_ptr(), out.data_ptr()
            arrays[k].row_floats, arrays[k].zero_new = row, int(zero_new)
            k += 1
        return out

    for name, t in params.items():
        new_params[name] = add(t, False)
    for name, (m, v) in state.items():
        assert m.shape == params[name].shape and m.is_contiguous() and v.is_contiguous()
        new_state[name] = (add(m, True), add(v, True))
    check(L.fg_refine_gather(n_out, n_keep, ptr(src), k, arrays, st))
    if n_children or n_kd:
        if samples is None:  # freegaussian_model.py:519 (rows of culled children are drawn too, as there)
            samples = torch.randn((samps * n_split, 3), device=dev, generator=generator)
        assert samples.shape == (samps * n_split, 3) and samples.is_cuda
        samples = samples.contiguous()
        check(L.fg_refine_children(n_out, n_keep, n_children, ptr(src), ptr(sample_row), ptr(samples), ptr(means),
                                   ptr(quats), ptr(scales), ptr(new_params["means"]), ptr(new_params["scales"]), st))
    return RefineResult(new_params, new_state, N, n_out, n_split, n_kd, n_keep, src, False)


def refine(params: Dict[str, Tensor], state: Dict[str, Tuple[Tensor, Tensor]], grad_norm: Optional[Tensor],
           vis_count: Optional[Tensor], max_size: Optional[Tensor], step: int, num_train_data: int,
           last_size: Tuple[int, int], cfg: RefineSchedule = RefineSchedule(), samples: Optional[Tensor] = None,
           generator: Optional[torch.Generator] = None) -> Optional[RefineResult]:
    """``refinement_after`` (``freegaussian_model.py:404-491``).  ``params`` holds the model's raw tensors
    (``scales`` log-space, ``opacities`` logit-space, any number of further per-Gaussian arrays such as
    ``features_dc`` / ``features_rest`` or the concatenated SH tensor); ``state[name] = (exp_avg, exp_avg_sq)``.
    Returns None before ``refine_start``; the caller drops its statistics afterwards (``:485-487``)."""
    if step < cfg.refine_start:
        return None
    reset_interval = cfg.reset_alpha_every * cfg.refine_every
    do_densification = step < cfg.stop_split_at and step % reset_interval > num_train_data + cfg.refine_every
    cull_big = step > cfg.refine_every * cfg.reset_alpha_every
    use_screen = step < cfg.stop_screen_size_at
    res: Optional[RefineResult]
    if do_densification:
        assert grad_norm is not None and vis_count is not None and max_size is not None
        res = plan_and_apply(params, state, grad_norm, vis_count, max_size, densify=True, use_screen=use_screen,
                             cull_big=cull_big, max_dim=float(max(last_size[0], last_size[1])), cfg=cfg,
                             samples=samples, generator=generator)
    elif step >= cfg.stop_split_at and cfg.continue_cull_post_densification:
        res = plan_and_apply(params, state, None, None, max_size, densify=False, use_screen=use_screen,
                             cull_big=cull_big, max_dim=1.0, cfg=cfg)
    else:
        n = int(params["means"].shape[0])
        res = RefineResult(params, state, n, n, 0, 0, n, None, False)
    if step < cfg.stop_split_at and step % reset_interval == cfg.refine_every:
        # opacity reset (:470-483): clamp the logits at logit(2 * cull_alpha_thresh), clear that group's moments
        reset_value = cfg.cull_alpha_thresh * 2.0
        cap = torch.logit(torch.tensor(reset_value)).item()
        res.params["opacities"].clamp_(max=cap)
        if "opacities" in res.state:
            m, v = res.state["opacities"]
            m.zero_()
            v.zero_()
        res.opacity_reset = True
    return res
