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

import os
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
    def sync(self, group=None, reduce: bool = True) -> None:
        """SUM, SUM, MAX across ranks of everything accumulated since the last sync, folded into the
        running statistics.  Call right before the statistics are read (refinement), or every step.
        ``reduce=False`` folds the local accumulators in without a collective (a rank that rendered every view)."""
        if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
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


class ViewShardedExchange:
    """The exchange step of a view-sharded run, executed INSIDE the projection backward over NVLink peer memory
    (``csrc/exchange.cu``).  After :meth:`install`, ``loss.backward()`` on every rank returns, for
    every Gaussian parameter of the ``rasterization`` call, the gradient summed over ALL ranks' views -- what a single
    process rendering every view would have computed.  Every rank must make the same sequence of
    ``rasterization(...)``/``backward()`` calls with the same N, views per rank and kwargs.

    What travels: the 12-byte colour gradient of every visible (view, Gaussian) is published in compact form in symmetric
    memory, every rank pulls the peers' published ranges over NVLink (coalesced 16-byte loads, ~3.5 MB per peer at 1 M
    Gaussians) and rebuilds the 192-byte SH rows of all views itself (``fg_xchg_sh_bwd_views``: fixed summation order,
    bit-identical on every rank); the geometry gradients (56 B per Gaussian: means, quats,
    scales, opacity, frame t+1 means) are summed in place by a two-shot all-reduce that runs in the NVSwitch
    (``fg_xchg_allreduce_f32``: ``multimem.ld_reduce`` + ``multimem.st``, barriers inside the kernel).  torch is used for
    the plumbing only: ``torch.distributed._symmetric_memory`` allocates and peer-maps the buffers.

    The gradient tensors handed to autograd are views of that symmetric buffer: they are overwritten by the next
    backward (set ``p.grad = None`` before every step, as the reference trainer's ``zero_grad(set_to_none=True)`` does).
    """

    def __init__(self, group=None, use_multicast: bool = True):
        assert dist.is_available() and dist.is_initialized(), "init_process_group first"
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.use_multicast = use_multicast and not os.environ.get("FG_XCHG_NO_MULTICAST")
        self.epoch = 0
        self.parity = 0
        self._buf = None       # uint8 tensor over this rank's symmetric block
        self._peers = None     # _lib.XchgPeers
        self._layout = None    # (pub_bytes, arena_bytes)
        self._old = []         # previous blocks are kept alive: a peer may still be reading them
        self._side = None      # side stream of the SH-row summation
        self._staging = None   # local copies of the peers' published blocks
        self._keep = None
        self.multicast = False

    # -- lifecycle
    def active(self) -> bool:
        # FG_XCHG_SOLO: run the exchange kernels in a single-rank group too (single-GPU tests and ncu captures of them)
        return self.world > 1 or bool(os.environ.get("FG_XCHG_SOLO"))

    def install(self) -> "ViewShardedExchange":
        from . import _lib, rendering
        for env, opt in (("FG_XCHG_AR_BLOCKS", b"xchg_ar_blocks"), ("FG_XCHG_PULL_BLOCKS", b"xchg_pull_blocks")):
            if os.environ.get(env):  # CTA counts of the communication kernels (sweeps; defaults 64 / 4)
                _lib.check(_lib.lib().fg_set_option(opt, int(os.environ[env])))
        rendering._exchange_hook = self
        return self

    def uninstall(self) -> None:
        from . import rendering
        if rendering._exchange_hook is self:
            rendering._exchange_hook = None

    # -- symmetric memory
    def _ensure(self, pub_bytes: int, arena_bytes: int, device) -> None:
        """(Re)allocate the symmetric block: flags | publish block x 2 (double-buffered by step parity) | arena.
        Collective: every rank computes the same sizes from the same N, so all ranks get here together."""
        from . import _lib
        import torch.distributed._symmetric_memory as symm
        cur = self._layout
        if cur is not None and pub_bytes <= cur[0] and arena_bytes <= cur[1]:
            return
        pub_bytes = max(pub_bytes, cur[0] if cur else 0)
        arena_bytes = max(arena_bytes, cur[1] if cur else 0)
        rnd = lambda b: (int(b) + 4095) // 4096 * 4096  # noqa: E731
        pub_bytes, arena_bytes = rnd(pub_bytes), rnd(arena_bytes)
        total = _lib.XCHG_FLAG_BYTES + 2 * pub_bytes + arena_bytes
        if self._buf is not None:
            torch.cuda.synchronize(device)
            self._old.append((self._buf, self._hdl))
        buf = symm.empty(total, dtype=torch.uint8, device=device)
        hdl = symm.rendezvous(buf, self.group.group_name)
        buf[:_lib.XCHG_FLAG_BYTES].zero_()
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every rank's flags are zero before anyone signals
        off = buf.data_ptr() - int(hdl.buffer_ptrs[hdl.rank])
        peers = _lib.XchgPeers()
        peers.world, peers.rank = self.world, self.rank
        for r in range(self.world):
            peers.buf[r] = int(hdl.buffer_ptrs[r]) + off
            peers.flags[r] = int(hdl.buffer_ptrs[r]) + off  # the flag area is the head of the block
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        # Bytes per link direction of the two-shot all-reduce of B bytes over G ranks: in the switch (multimem) every
        # rank's whole buffer is read once and every rank receives the whole result, (G+1)/G x B -- its own slice loops
        # through the switch too; with peer loads + peer stores 2(G-1)/G x B.  The switch wins from G = 4 on (G = 2: 1.5 B
        # against 1.0 B; measured 0.185 ms against the peer path for 68 MB).
        self.multicast = bool(mc) and self.use_multicast and (self.world >= 4 or bool(os.environ.get("FG_XCHG_FORCE_MULTICAST")))
        peers.mc = (mc + off) if self.multicast else None
        self._buf, self._hdl, self._peers = buf, hdl, peers
        self._layout = (pub_bytes, arena_bytes)
        self.pub_off = [_lib.XCHG_FLAG_BYTES, _lib.XCHG_FLAG_BYTES + pub_bytes]
        self.arena_off = _lib.XCHG_FLAG_BYTES + 2 * pub_bytes

    def prepare(self, n_floats: int, C: int, N: int, device) -> None:
        """Size the symmetric block for a backward with an ``n_floats`` arena and C published views of N Gaussians
        (collective the first time and whenever it has to grow)."""
        from . import _lib
        pub = int(_lib.lib().fg_xchg_pub_bytes(C, N)) if C else 0
        self._ensure(pub, (n_floats + 3) // 4 * 16, device)

    def arena(self, n_floats: int, device) -> Tensor:
        """The flat gradient buffer of one backward: ``n_floats`` fp32 of symmetric memory."""
        self._ensure(0, (n_floats + 3) // 4 * 16, device)
        return self._buf[self.arena_off:self.arena_off + 4 * n_floats].view(torch.float32)

    def publish_block(self, C: int, N: int, offsets: Tensor, nnz_dev: Tensor):
        """``fg_project_bwd_pub`` pointing at this step's publish block (camera centres + nnz | mask | prefix | compact
        colour gradients); ``offsets`` / ``nnz_dev`` = the exclusive scan of the visibility (``fg_pack_plan``)."""
        import ctypes as C_
        from . import _lib
        L = _lib.lib()
        assert self._layout is not None and int(L.fg_xchg_pub_bytes(C, N)) <= self._layout[0], "prepare() first"
        o = [C_.c_int64() for _ in range(4)]
        words = C_.c_int32()
        _lib.check(L.fg_xchg_pub_layout(C, N, *[C_.byref(x) for x in o], C_.byref(words)))
        base = self._buf.data_ptr() + self.pub_off[self.parity]
        pub = _lib.ProjectBwdPub()
        pub.campos, pub.mask, pub.prefix, pub.rgb = base, base + o[1].value, base + o[2].value, base + o[3].value
        pub.offsets, pub.nnz, pub.words = offsets.data_ptr(), nnz_dev.data_ptr(), words.value
        self._keep = (offsets, nnz_dev)  # alive until the kernels that read them have been enqueued behind them
        return pub

    def side_stream(self, device):
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def barrier_published(self) -> None:
        """After the publishing SH kernel of this step, on the current stream: cross-rank barrier (every rank has
        published); the event recorded behind it is what the side stream waits for."""
        from . import _lib
        cur = torch.cuda.current_stream()
        self.epoch += 1
        _lib.check(_lib.lib().fg_xchg_barrier(self._peers, self.epoch, cur.cuda_stream))
        self._published = torch.cuda.Event()
        self._published.record(cur)

    def sh_rows_async(self, C: int, N: int, sh_degree, sh_bases: int, means: Tensor, v_sh: Tensor) -> None:
        """On a side stream, behind :meth:`barrier_published`: pull the peers' published ranges and rebuild the SH rows
        from every rank's views (``fg_xchg_sh_bwd_views``) while the current stream runs the geometry kernel and its
        all-reduce (enqueueing the geometry kernel first was measured too: 2.06 ms per step at 8 ranks against 1.95 ms).
        :meth:`join` makes the current stream wait for the rows."""
        from . import _lib
        from .rendering import _stage
        L = _lib.lib()
        self.side_stream(means.device)
        stride = int(L.fg_xchg_pub_bytes(C, N))
        need = stride * self.world if self.world > 1 else 0
        if self._staging is None or self._staging.numel() < need:
            self._staging = torch.empty(max(need, 16), dtype=torch.uint8, device=means.device)  # pulled copies: local memory
        self._side.wait_event(self._published)
        with torch.cuda.stream(self._side):
            with _stage("xchg_sh_views"):
                _lib.check(L.fg_xchg_sh_bwd_views(self._peers, self.pub_off[self.parity], C, N, int(sh_degree), sh_bases,
                                                  _lib.ptr(means), _lib.ptr(v_sh), _lib.ptr(self._staging) if need else None,
                                                  stride, self._side.cuda_stream))
        self.parity ^= 1

    def join(self) -> None:
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    def reduce(self, n_floats: int) -> None:
        """In-place all-reduce(SUM) of the first ``n_floats`` of the arena (two-shot, in the switch when multicast is
        available).  Enqueued on the current stream; no host synchronisation."""
        from . import _lib
        from .rendering import _stage
        self.epoch += 1
        n4 = (n_floats + 3) // 4 * 4
        with _stage("xchg_allreduce"):
            _lib.check(_lib.lib().fg_xchg_allreduce_f32(self._peers, self.arena_off, n4, self.epoch, 1,
                                                       torch.cuda.current_stream().cuda_stream))

    def all_reduce_(self, t: Tensor) -> Tensor:
        """In-place SUM of any fp32 CUDA tensor over the ranks through the same NVLS kernel (staged through the arena):
        network weight gradients, densification statistics."""
        from . import _lib
        flat = t.reshape(-1)
        n = flat.numel()
        a = self.arena(n, t.device)
        a[:n].copy_(flat)
        self.epoch += 1
        _lib.check(_lib.lib().fg_xchg_allreduce_f32(self._peers, self.arena_off, (n + 3) // 4 * 4, self.epoch, 1,
                                                   torch.cuda.current_stream().cuda_stream))
        flat.copy_(a[:n])
        return t


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
    hook = rendering._exchange_hook
    if hook is not None and hook.active() and hook._buf is not None:
        # gradients that came out of a backward with ViewShardedExchange installed are already summed over the ranks
        base = hook._buf.untyped_storage().data_ptr()
        grads = [g for g in grads if g.untyped_storage().data_ptr() != base]
        if not grads:
            return
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
