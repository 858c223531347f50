"""View-sharded multi-GPU training support (BASELINE.json north_star item 5).

One process per GPU.  Gaussian parameters are replicated; rank r renders the views
``{c : c mod G == r}`` of the step's batch with no forward communication (cameras are
independent: the camera id is the top field of the sort key).  After backward there is ONE
exchange step over NCCL/NVLink:

* all-reduce(SUM) of the Gaussian-parameter gradients, coalesced into one flat fp32 buffer
  (236 B per Gaussian: means 3 + quats 4 + scales 3 + opacity 1 + SH 48 floats);
* all-reduce(SUM) of ``xys_grad_norm`` and ``vis_counts`` and all-reduce(MAX) of
  ``max_2Dsize`` -- the densification statistics of ``freegaussian_model.py:369-392``.

The reference has no multi-GPU code; the statistics arithmetic below is the
reference's ``after_train_iter`` restated per view so that the reduced result equals what a
single process rendering all views would accumulate.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(n_views: int, rank: int, world_size: int) -> List[int]:
    """Indices of the views rank ``rank`` renders: ``c mod G == r``."""
    assert 0 <= rank < world_size
    return list(range(rank, n_views, world_size))


class GradBucket:
    """One flat fp32 buffer holding every Gaussian-parameter gradient, so the exchange is a
    single NCCL all-reduce sized for launch latency rather than one per tensor."""

    def __init__(self, params: Sequence[Tensor]):
        self.shapes = [p.shape for p in params]
        self.numels = [p.numel() for p in params]
        total = sum(self.numels)
        self.flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        self.views = []
        o = 0
        for p, n in zip(params, self.numels):
            self.views.append(self.flat[o : o + n].view(p.shape))
            o += n

    def attach(self, params: Sequence[Tensor]) -> None:
        """Point each ``p.grad`` at its slice so backward accumulates straight into the bucket."""
        for p, v in zip(params, self.views):
            p.grad = v

    def zero_(self) -> None:
        self.flat.zero_()

    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def all_reduce(self, group=None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class DensificationStats:
    """``xys_grad_norm`` / ``vis_counts`` / ``max_2Dsize`` of ``freegaussian_model.py:369-392``.

    Every step costs one fused kernel (``fg_densify_stats``) that folds the rank's views into per-rank
    accumulators.  All three statistics are plain accumulators (SUM, SUM, MAX), so the cross-rank
    reduction is only needed when they are consumed -- the refinement step every ``refine_every`` = 100
    iterations (``config/sim/base.yaml:22``): :meth:`sync` all-reduces what was accumulated since the last
    sync and folds it into the public running values, which then equal what a single process rendering
    every view (or a per-step reduction) would hold."""

    def __init__(self, n: int, device):
        self.xys_grad_norm = torch.zeros(n, dtype=torch.float32, device=device)
        self.vis_counts = torch.ones(n, dtype=torch.float32, device=device)  # reference starts at one (:380)
        self.max_2Dsize = torch.zeros(n, dtype=torch.float32, device=device)
        self._local_grad = torch.zeros_like(self.xys_grad_norm)
        self._local_vis = torch.zeros_like(self.vis_counts)
        self._local_size = torch.zeros_like(self.max_2Dsize)

    @torch.no_grad()
    def accumulate_local(self, radii: Tensor, absgrad: Tensor, height: int, width: int) -> None:
        """Fold this rank's views in: radii [C,N] int32, absgrad [C,N,2] (``meta["means2d"].absgrad``)."""
        if radii.is_cuda:  # one fused kernel (fg_densify_stats) instead of ~10 elementwise launches
            from . import _lib
            C, N = radii.shape
            _lib.check(_lib.lib().fg_densify_stats(
                C, N, _lib.ptr(radii.contiguous()), _lib.ptr(absgrad.contiguous()), 1.0 / float(max(height, width)),
                _lib.ptr(self._local_grad), _lib.ptr(self._local_vis), _lib.ptr(self._local_size),
                torch.cuda.current_stream().cuda_stream))
            return
        # host tensors (gloo tests of the exchange logic): the same arithmetic in torch
        vis = radii > 0
        norms = absgrad.norm(dim=-1)
        self._local_grad += torch.where(vis, norms, torch.zeros_like(norms)).sum(0)
        self._local_vis += vis.sum(0).to(torch.float32)
        size = radii.to(torch.float32) / float(max(height, width))
        self._local_size = torch.maximum(self._local_size, size.max(0).values)

    @torch.no_grad()
    def sync(self, group=None) -> None:
        """SUM, SUM, MAX across ranks of everything accumulated since the last sync, folded into the
        running statistics.  Call right before the statistics are read (refinement), or every step."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            packed = torch.stack([self._local_grad, self._local_vis])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            self._local_grad, self._local_vis = packed[0], packed[1]
            dist.all_reduce(self._local_size, op=dist.ReduceOp.MAX, group=group)
        self.xys_grad_norm += self._local_grad
        self.vis_counts += self._local_vis
        self.max_2Dsize = torch.maximum(self.max_2Dsize, self._local_size)
        self._local_grad = torch.zeros_like(self._local_grad)
        self._local_vis = torch.zeros_like(self._local_vis)
        self._local_size = torch.zeros_like(self._local_size)

    reduce = sync  # per-step use


def exchange(grads: Sequence[Tensor], group=None) -> None:
    """THE exchange step of a view-sharded training step: all-reduce(SUM) of every
    parameter gradient.  (The densification statistics are reduced by ``DensificationStats.sync``.)

    The projection backward writes all of its parameter gradients into one flat buffer
    (``rendering.last_grad_arena()``); gradients living there are reduced by ONE collective over that
    buffer, and small leftovers (the opacity gradient) ride in its spare tail.  Separate collectives per
    tensor cost ~0.25 ms of launch latency per step on top of the ~0.5 ms the 236 MB need on NVLink at
    1 M Gaussians."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    from . import rendering
    grads = [g for g in grads if g is not None]
    info = rendering.last_grad_arena() if grads and grads[0].is_cuda else None
    loose = grads
    if info is not None:
        arena, used = info
        base = arena.untyped_storage().data_ptr()
        inside = [g for g in grads if g.untyped_storage().data_ptr() == base]
        loose = [g for g in grads if g.untyped_storage().data_ptr() != base]
        if inside:
            small = [t for t in loose if t.is_contiguous()]
            n_small = sum(t.numel() for t in small)
            if n_small <= arena.numel() - used:
                tail = arena[used:used + n_small]
                if small:
                    torch.cat([t.reshape(-1) for t in small], out=tail)
                dist.all_reduce(arena[:used + n_small], op=dist.ReduceOp.SUM, group=group)
                o = 0
                for t in small:
                    t.copy_(tail[o:o + t.numel()].view_as(t))
                    o += t.numel()
                loose = [t for t in loose if not t.is_contiguous()]
            else:
                dist.all_reduce(arena[:used], op=dist.ReduceOp.SUM, group=group)
    works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in loose]
    for w in works:
        w.wait()
