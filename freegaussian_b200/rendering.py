"""``rasterization(...)`` -- the drop-in boundary of the B200-native splat renderer.

Same call signature, return tuple and ``meta`` keys as ``gsplat.rendering.rasterization``
(gsplat 1.4 semantics) as FreeGaussian calls it at
``freegaussian/freegaussian_model.py:847-868``, ``freegaussian_control_model.py:158-179``
and (``packed=True``, ``"ED"``) ``preprocess/knn_gaussian.py:93-113``.  Host logic only:
every arithmetic step runs in the hand-written sm_100a kernels behind the C ABI of
``include/fg_api.h`` (``libfreegaussian_b200.so``), on ``torch.cuda.current_stream()``.
There is no CPU path: non-CUDA inputs raise.

Extension (BASELINE.json north_star): pass ``means_next`` (frame t+1
means; optionally ``quats_next`` / ``scales_next`` with ``flow_mode="cov"``) and the rendered
Gaussian flow ``sum_i T_i alpha_i (mu2d_i(t+1) - mu2d_i(t))`` (``docs/index.html:286-299``)
comes back as ``meta["flow"]`` ``[C,H,W,2]``, composited in the same pass as RGB and depth.
Without these kwargs the behaviour is exactly the reference call's.
"""

from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr

MAX_CH = 8  # FG_MAX_CHANNELS
FUSED_CALLS = True  # forward pass through fg_render_front / fg_render_back (two C calls) when possible
SORT_MODE = "binned"  # "binned" | "two_level" | "key64" (the reference's literal 64-bit key sort); same lists
BIN_RANKED = True  # granular "binned" mode: ranked placement of the coarse pairs when it applies (False: emit + sort, for tests)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Workspace:
    """Per-device scratch buffers reused across calls (safe to drop at any time)."""

    def __init__(self):
        self.bufs: Dict[Tuple[str, int], Tensor] = {}

    def get(self, name: str, nbytes: int, device) -> Tensor:
        key = (name, torch.device(device).index or 0)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            # a buffer that has to grow grows by a quarter beyond the need: a size that creeps upward call after call
            # (list lengths of a scene in training) must not turn into a cudaMalloc per call
            if buf is not None:
                nbytes = _round_up(nbytes + nbytes // 4, 1 << 20)
            self.bufs[key] = buf = None  # the old buffer goes back to the allocator before the new one is asked for
            buf = torch.empty(max(nbytes, 1024), dtype=torch.uint8, device=device)
            self.bufs[key] = buf
        return buf


def _round_up(x: int, q: int) -> int:
    return (x + q - 1) // q * q


_Counts3 = ctypes.c_int64 * 3  # (M, Mc, lists built) written by fg_render_front
_LIST_QUANTUM = 1 << 22  # entries (16 MiB of int32): list capacities are multiples of this


def _list_capacity(cur: int, need: int) -> int:
    """Capacity of a guessed-size list buffer for the next call, given the current capacity and the last call's need.

    Stable by construction: the capacity changes only when the need comes within 8 % of it, and then jumps to 1.3 x the
    need rounded up to 16 MiB, so that a need that drifts (Gaussians move while training) changes the buffer size -- and
    sends torch's caching allocator to cudaMalloc, 10 - 200 ms with kernels in flight -- once per ~20 % of growth instead
    of at every new maximum (bench.py's train_iter section, round 2: 20 - 29 cudaMallocs per 90 iterations before this).
    """
    if need > cur - cur // 12:
        return _round_up(need + (3 * need) // 10 + 1024, _LIST_QUANTUM)
    return cur


_ws = _Workspace()

# The flat buffer holding the parameter gradients of the most recent projection backward:
# (tensor, floats in use).  Read by freegaussian_b200.dist.exchange.
_grad_arena: Dict[str, tuple] = {}
_list_guess: Dict[tuple, tuple] = {}  # (C, N, W, H, device) -> (capacity of flatten_ids, of the coarse pairs)
_list_small: Dict[tuple, int] = {}    # consecutive calls that needed less than half of that capacity
STATS = {"list_capacity_changes": 0, "list_second_call": 0, "list_len_min": 0, "list_len_max": 0}  # diagnostics (bench.py)


def last_grad_arena():
    return _grad_arena.get("last")


# Installed by freegaussian_b200.dist.ViewShardedExchange.install(): when set (and world size > 1) the projection
# backward writes into symmetric memory and ends with the cross-rank exchange, so the gradients autograd hands to the
# parameters are already the sums over every rank's views.
_exchange_hook = None


def _align4(n: int) -> int:
    return (n + 3) & ~3


class StageTimer:
    """Optional CUDA-event timing of each C-ABI stage (bench.py's per-kernel roofline).
    Disabled by default: ``with _stage(name)`` then costs one attribute test."""

    def __init__(self):
        self.enabled = False
        self.events = []  # (name, start, end)

    def reset(self):
        self.events = []

    def summary(self):
        torch.cuda.synchronize()
        out: Dict[str, list] = {}
        for name, s, e in self.events:
            out.setdefault(name, []).append(s.elapsed_time(e))
        return out


stage_timer = StageTimer()


class _stage:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if stage_timer.enabled:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if stage_timer.enabled:
            self.e.record()
            stage_timer.events.append((self.name, self.s, self.e))
        return False


def _require_cuda(**tensors) -> None:
    for name, t in tensors.items():
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                f"freegaussian_b200.rasterization: `{name}` is on {t.device}; the renderer has no CPU path "
                "(the CUDA kernels are the only implementation)"
            )
        assert t.dtype == torch.float32, f"{name} must be float32, got {t.dtype}"


# --------------------------------------------------------------------------- projection
class _Project(torch.autograd.Function):
    """fg_project_fwd / fg_project_bwd.  Outputs: radii, means2d, depths, conics, comps, feat, tiles."""

    @staticmethod
    def forward(ctx, means, quats, scales, colors, means_next, quats_next, scales_next, viewmats, Ks, cfg,
                opacities=None):
        L = _lib.lib()
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        means, quats, scales = means.contiguous(), quats.contiguous(), scales.contiguous()
        viewmats, Ks = viewmats.contiguous(), Ks.contiguous()
        sh_degree = cfg["sh_degree"]
        use_sh = sh_degree is not None
        if use_sh:
            colors = colors.contiguous()
            sh_bases = colors.shape[-2]
            n_col = 3
        else:
            sh_bases = 0
            n_col = 0 if colors is None else colors.shape[-1]
        want_depth, want_flow = cfg["want_depth"], means_next is not None
        if means_next is not None:
            means_next = means_next.contiguous()
        flow_cov = bool(cfg.get("flow_cov")) and want_flow
        quats_next = quats_next.contiguous() if (flow_cov and quats_next is not None) else None
        scales_next = scales_next.contiguous() if (flow_cov and scales_next is not None) else None
        CH = n_col + (1 if want_depth else 0) + (2 if want_flow else 0)
        rgb_off = 0 if use_sh else -1
        depth_off = n_col if want_depth else -1
        flow_off = n_col + (1 if want_depth else 0) if want_flow else -1

        radii = torch.empty(C, N, dtype=torch.int32, device=dev)
        means2d = torch.empty(C, N, 2, device=dev)
        depths = torch.empty(C, N, device=dev)
        conics = torch.empty(C, N, 3, device=dev)
        comps = torch.empty(C, N, device=dev) if cfg["antialiased"] else None
        feat = torch.empty(C, N, CH, device=dev)
        tiles = torch.empty(C, N, dtype=torch.int32, device=dev)
        flow_affine = torch.empty(C, N, 4, device=dev) if flow_cov else None
        if cfg.get("fused"):
            # projection + depth sort + binning up to the host sync, in one C call
            tile_w = math.ceil(cfg["width"] / cfg["tile_size"])
            tile_h = math.ceil(cfg["height"] / cfg["tile_size"])
            order = torch.empty(C * N, dtype=torch.int32, device=dev)
            coarse_off = torch.empty(C * N, dtype=torch.int32, device=dev)
            isect_offsets = torch.empty(C, tile_h, tile_w, dtype=torch.int32, device=dev)
            ws = _ws.get("front", L.fg_render_front_workspace_bytes(C, N, tile_w, tile_h), dev)
            # the list length M is only known after the call's host sync; a buffer of a guessed capacity
            # (1.25 x the largest M seen for this shape) lets the C side go on to build the lists without
            # coming back here first.  A wrong guess costs one extra call, never a wrong result.
            key = (C, N, cfg["width"], cfg["height"], dev)
            guess_m, guess_mc = _list_guess.get(key, (0, 0))
            flat_buf = torch.empty(guess_m, dtype=torch.int32, device=dev) if guess_m else None
            ws2 = _ws.get("back", L.fg_render_back_workspace_bytes(C, tile_w, tile_h, guess_mc), dev) if guess_m else None
            counts = _Counts3()
            check(L.fg_render_front(
                C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), cfg["width"], cfg["height"],
                cfg["eps2d"], cfg["near_plane"], cfg["far_plane"], cfg["radius_clip"], cfg["tile_size"],
                sh_degree if use_sh else -1, sh_bases, ptr(colors) if use_sh else None, ptr(means_next),
                ptr(quats_next), ptr(scales_next), int(flow_cov), ptr(radii), ptr(means2d), ptr(depths), ptr(conics),
                ptr(comps), ptr(feat), CH, rgb_off, depth_off, flow_off, ptr(flow_affine), ptr(tiles), ptr(order),
                ptr(isect_offsets), ptr(coarse_off), counts, ptr(ws), ws.numel(), ptr(flat_buf), guess_m,
                ptr(ws2), ws2.numel() if ws2 is not None else 0, _stream()))
            M, Mc = int(counts[0]), int(counts[1])
            # capacity for the next call (_list_capacity: stable sizes, so that the allocator hands back the same blocks
            # call after call); after 64 consecutive calls that needed less than a third of it the capacity is re-derived
            # from the current need (a scene that shrank -- culling after densification -- gives its buffers back)
            small = _list_small.get(key, 0) + 1 if (3 * M + 3 * _LIST_QUANTUM < guess_m) else 0
            if small >= 64:
                guess_m, guess_mc, small = 0, 0, 0
            _list_small[key] = small
            new_guess = (_list_capacity(guess_m, M), _list_capacity(guess_mc, Mc))
            STATS["list_capacity_changes"] += new_guess != _list_guess.get(key)
            STATS["list_second_call"] += not counts[2]
            STATS["list_len_min"] = min(STATS["list_len_min"] or M, M)
            STATS["list_len_max"] = max(STATS["list_len_max"], M)
            _list_guess[key] = new_guess
            if counts[2]:
                flatten_ids = flat_buf[:M]
            else:
                flatten_ids = torch.empty(M, dtype=torch.int32, device=dev)
                ws2 = _ws.get("back", L.fg_render_back_workspace_bytes(C, tile_w, tile_h, Mc), dev)
                check(L.fg_render_back(C, N, M, Mc, ptr(order), ptr(coarse_off), ptr(means2d), ptr(radii),
                                       cfg["tile_size"], ptr(isect_offsets), ptr(flatten_ids), ptr(ws2), ws2.numel(),
                                       ptr(ws), ws.numel(), 0,
                                       cfg["width"], cfg["height"], None, None, None, None, None, -1, 0, -1, 0, None,
                                       None, None, None, _stream()))
            cfg["_front"] = dict(isect_offsets=isect_offsets, flatten_ids=flatten_ids, M=M, Mc=Mc, tile_w=tile_w,
                                 tile_h=tile_h)
        else:
            with _stage("project_fwd"):
                check(L.fg_project_fwd(
                    C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), cfg["width"], cfg["height"],
                    cfg["eps2d"], cfg["near_plane"], cfg["far_plane"], cfg["radius_clip"], cfg["tile_size"],
                    sh_degree if use_sh else -1, sh_bases, ptr(colors) if use_sh else None, ptr(means_next),
                    ptr(quats_next), ptr(scales_next), int(flow_cov), ptr(radii), ptr(means2d), ptr(depths),
                    ptr(conics), ptr(comps), ptr(feat), CH, rgb_off, depth_off, flow_off, ptr(flow_affine),
                    ptr(tiles), _stream()))
        if not use_sh and colors is not None:
            feat[..., :n_col] = colors if colors.dim() == 3 else colors[None]
        # view-sharded run: the backward publishes the colour gradients of the visible (view, Gaussian) pairs in compact
        # form; their rows come from an exclusive scan of the visibility, done here where the radii are produced
        ctx.pub_scan = None
        xc = _exchange_hook
        if xc is not None and xc.active() and use_sh and sh_bases % 4 == 0 and 0 < N and C <= 128:
            offs = torch.empty(C * N, dtype=torch.int32, device=dev)
            nnz_dev = torch.empty(1, dtype=torch.int64, device=dev)
            ws_scan = _ws.get("pub_scan", L.fg_pack_workspace_bytes(C * N), dev)
            # on the exchange's side stream: nothing of the forward pass waits for it, the backward does
            side = xc.side_stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            check(L.fg_pack_plan(C * N, ptr(radii), ptr(offs), ptr(nnz_dev), ptr(ws_scan), ws_scan.numel(), side.cuda_stream))
            done = torch.cuda.Event()
            done.record(side)
            ctx.pub_scan = (offs, nnz_dev, done)
        ctx.save_for_backward(means, quats, scales, colors if use_sh else None, means_next, quats_next, scales_next,
                              viewmats, Ks, radii, feat if use_sh else None)
        ctx.cfg = cfg
        ctx.layout = (CH, n_col, rgb_off, depth_off, flow_off, sh_bases, use_sh,
                      None if colors is None else colors.dim(), flow_cov)
        ctx.mark_non_differentiable(radii, tiles)
        if comps is None:
            comps = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(comps)
        if flow_affine is None:
            flow_affine = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(flow_affine)
        # the opacities pass through unchanged: their gradient then arrives in THIS node's backward together with every
        # other parameter gradient of the call, which is where a view-sharded run sums them across ranks
        opac_out = torch.empty(0, device=dev) if opacities is None else opacities.view(opacities.shape)
        if opacities is None:
            ctx.mark_non_differentiable(opac_out)
        return radii, means2d, depths, conics, comps, feat, tiles, flow_affine, opac_out

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, v_comps, v_feat, _v_tiles, v_flow_affine, v_opac):
        L = _lib.lib()
        means, quats, scales, sh, means_next, quats_next, scales_next, viewmats, Ks, radii, feat_fwd = ctx.saved_tensors
        cfg = ctx.cfg
        CH, n_col, rgb_off, depth_off, flow_off, sh_bases, use_sh, col_dim, flow_cov = ctx.layout
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device

        def c(t):
            return None if t is None else t.contiguous()

        v_means2d, v_depths, v_conics, v_feat = c(v_means2d), c(v_depths), c(v_conics), c(v_feat)
        v_comps = c(v_comps) if cfg["antialiased"] else None
        xc = _exchange_hook if (_exchange_hook is not None and _exchange_hook.active()) else None
        v_opac = v_opac if ctx.needs_input_grad[10] else None
        v_col2d = None  # colours given per Gaussian ([N,D], shared by the cameras) instead of SH coefficients
        if not use_sh and col_dim == 2 and v_feat is not None:
            v_col2d = v_feat[..., :n_col].sum(0)
        # All parameter gradients of this call live in ONE flat buffer: the geometry segments first (each padded to
        # 16 bytes; in a view-sharded run also the gradients that only pass through this node), the SH rows last.
        # Single GPU: a fresh allocation.  View-sharded: symmetric memory, reduced in place below.
        geo = [3 * N, 4 * N, 3 * N, 3 * N if means_next is not None else 0, 4 * N if quats_next is not None else 0,
               3 * N if scales_next is not None else 0,
               v_opac.numel() if (xc is not None and v_opac is not None) else 0,
               v_col2d.numel() if (xc is not None and v_col2d is not None) else 0]
        offs, o = [], 0
        for n_ in geo:
            offs.append(o)
            o += _align4(n_)
        geo_floats = o
        sh_floats = sh_bases * 3 * N if use_sh else 0
        used = geo_floats + sh_floats
        # view-sharded with SH colours: publish the 12-byte colour gradients instead of writing 192-byte SH rows; the rows
        # are then summed over every rank's views by fg_xchg_sh_bwd_views and only the geometry segments are all-reduced
        want_pub = xc is not None and ctx.pub_scan is not None and v_feat is not None
        if xc is not None:
            xc.prepare(used, C if want_pub else 0, N, dev)
        arena = xc.arena(used, dev) if xc is not None else torch.empty(used + 3 * N + 4, device=dev)
        seg = lambda k, shape: arena[offs[k]:offs[k] + geo[k]].view(shape) if geo[k] else None  # noqa: E731
        v_means, v_quats, v_scales = seg(0, (N, 3)), seg(1, (N, 4)), seg(2, (N, 3))
        v_means_next, v_quats_next, v_scales_next = seg(3, (N, 3)), seg(4, (N, 4)), seg(5, (N, 3))
        v_sh = arena[geo_floats:used].view(N, sh_bases, 3) if use_sh else None
        _grad_arena["last"] = (arena, used)
        v_flow_affine = c(v_flow_affine) if flow_cov else None
        pub = None
        if want_pub:
            torch.cuda.current_stream().wait_event(ctx.pub_scan[2])  # the visibility scan of the forward pass
            pub = xc.publish_block(C, N, ctx.pub_scan[0], ctx.pub_scan[1])

        def project_bwd(phase):
            if pub is not None:
                pub.phase = phase
            check(L.fg_project_bwd(
                C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), cfg["width"], cfg["height"],
                cfg["eps2d"], cfg["near_plane"], cfg["far_plane"], cfg["radius_clip"],
                cfg["sh_degree"] if use_sh else -1, sh_bases, ptr(sh), ptr(means_next), ptr(quats_next), ptr(scales_next),
                int(flow_cov), ptr(radii), ptr(v_means2d), ptr(v_depths), ptr(v_conics), ptr(v_comps), ptr(v_feat),
                ptr(feat_fwd), CH, rgb_off, depth_off, flow_off, ptr(v_flow_affine), ptr(v_means), ptr(v_quats), ptr(v_scales),
                None if pub is not None else ptr(v_sh), ptr(v_means_next), ptr(v_quats_next), ptr(v_scales_next),
                None if pub is None else ctypes.byref(pub), _stream()))

        if pub is None:
            with _stage("project_bwd"):
                project_bwd(0)
        if xc is not None:
            if geo[6]:
                seg(6, v_opac.shape).copy_(v_opac)
                v_opac = seg(6, v_opac.shape)
            if geo[7]:
                seg(7, v_col2d.shape).copy_(v_col2d)
                v_col2d = seg(7, v_col2d.shape)
            if pub is not None:
                # SH kernel (publishes) -> barrier -> [side stream: SH rows from every rank's views] || [geometry kernel ->
                # in-switch all-reduce]; the side stream is joined before the gradients are handed to autograd
                with _stage("project_bwd"):
                    project_bwd(1)
                with _stage("exchange"):
                    xc.barrier_published()
                    xc.sh_rows_async(C, N, cfg["sh_degree"], sh_bases, means, v_sh)
                    with _stage("project_bwd_geo"):
                        project_bwd(2)
                    xc.reduce(geo_floats)
                    xc.join()
            else:
                with _stage("exchange"):
                    xc.reduce(used)
        v_colors = None
        if use_sh:
            v_colors = v_sh
        elif col_dim is not None and v_feat is not None:
            v_colors = v_col2d if col_dim == 2 else v_feat[..., :n_col]
        return (v_means, v_quats, v_scales, v_colors, v_means_next, v_quats_next, v_scales_next, None, None, None,
                v_opac)


# --------------------------------------------------------------------------- tile intersection
def _sort_pairs(L, n, keys_a, vals_a, keys_b, vals_b, end_bit, dev, st, u64):
    ws = _ws.get("sort", L.fg_radix_sort_workspace_bytes(n), dev)
    sel = ctypes.c_int(0)
    fn = L.fg_radix_sort_pairs_u64_u32 if u64 else L.fg_radix_sort_pairs_u32_u32
    check(fn(n, ptr(keys_a), ptr(vals_a), ptr(keys_b), ptr(vals_b), end_bit, ptr(ws), ws.numel(), ctypes.byref(sel), st))
    return (keys_b, vals_b) if sel.value == 1 else (keys_a, vals_a)


@torch.no_grad()
def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tiles_per_gauss: Tensor, tile_size: int,
                tile_w: int, tile_h: int, mode: str = "binned"):
    """Tile intersections sorted by (camera, tile, depth), ties in ascending c*N+n -- the order
    gsplat's 64-bit stable radix sort produces.

    ``mode="binned"`` (default): no sort over the intersections at all -- splats are sorted by depth
    once, exact per-tile counts come from a 2-D difference grid, and each 4x4-tile cell appends
    its depth-ordered splats to its tiles' lists (csrc/binning.cu).
    ``mode="key64"``: the reference layout literally -- emit 64-bit (camera|tile|depth) keys in
    (c,n) order and radix-sort them (6-7 passes).  ``mode="two_level"``: sort the splats once by
    depth, emit their tiles in that order with 32-bit tile keys and stable-sort by tile (2 passes).
    All three produce the identical lists (tests/test_gpu_stages.py).

    Returns ``(isect_ids | None, flatten_ids [M] int32, isect_offsets [C,tile_h,tile_w] int32,
    tile_keys | None)``; one host sync (reading M), like gsplat.
    """
    assert mode in ("binned", "two_level", "key64"), mode
    L = _lib.lib()
    C, N = radii.shape
    dev = radii.device
    st = _stream()
    total = C * N
    i32 = dict(dtype=torch.int32, device=dev)
    isect_offsets = torch.empty(C, tile_h, tile_w, **i32)
    n_dev = torch.empty(1, dtype=torch.int64, device=dev)
    ws = _ws.get("scan", L.fg_scan_workspace_bytes(total), dev)
    offsets = torch.empty(total, **i32)

    if mode == "key64":
        with _stage("scan"):
            check(L.fg_exclusive_scan_i32(total, ptr(tiles_per_gauss), ptr(offsets), ptr(n_dev), ptr(ws), ws.numel(), st))
        M = int(n_dev.item())  # host sync: sizes the intersection buffers
        assert M < 2**31, "too many tile intersections"
        ids_a = torch.empty(M, dtype=torch.int64, device=dev)
        val_a = torch.empty(M, **i32)
        if M > 0:
            with _stage("emit"):
                check(L.fg_isect_emit(C, N, ptr(means2d), ptr(radii), ptr(depths), ptr(offsets), tile_size, tile_w,
                                      tile_h, ptr(ids_a), ptr(val_a), st))
            tile_bits = int(math.floor(math.log2(tile_w * tile_h))) + 1
            cam_bits = int(math.floor(math.log2(C))) + 1
            with _stage("sort"):
                ids_a, val_a = _sort_pairs(L, M, ids_a, val_a, torch.empty_like(ids_a), torch.empty_like(val_a),
                                           32 + tile_bits + cam_bits, dev, st, True)
        with _stage("offsets"):
            check(L.fg_isect_offsets(M, ptr(ids_a), C, tile_w, tile_h, ptr(isect_offsets), st))
        return ids_a, val_a, isect_offsets, None

    # ---- both remaining modes start from the splats sorted by depth
    with _stage("depth_sort"):
        if mode == "binned":  # visible splats only: keys + histograms in one kernel, culled splats dropped by the first pass
            order = torch.empty(total, **i32)
            n_vis_dev = torch.empty(1, dtype=torch.int64, device=dev)
            wsd = _ws.get("depth_sort", L.fg_depth_sort_workspace_bytes(total), dev)
            check(L.fg_depth_sort_visible(total, ptr(depths), ptr(tiles_per_gauss), ptr(order), ptr(n_vis_dev), ptr(wsd),
                                          wsd.numel(), st))
        else:
            dk, dv = torch.empty(total, **i32), torch.empty(total, **i32)
            check(L.fg_isect_depth_keys(total, ptr(depths), ptr(tiles_per_gauss), ptr(dk), ptr(dv), st))
            _, order = _sort_pairs(L, total, dk, dv, torch.empty_like(dk), torch.empty_like(dv), 32, dev, st, False)

    if mode == "binned":
        cw_, ch_ = ctypes.c_int(0), ctypes.c_int(0)
        check(L.fg_bin_coarse_dims(tile_w, tile_h, ctypes.byref(cw_), ctypes.byref(ch_)))
        cw, chh = cw_.value, ch_.value
        n2 = torch.empty(2, dtype=torch.int64, device=dev)
        # few coarse cells (one or two 1080p views): the (splat, cell) pairs are placed by rank, never sorted
        rank_bytes = L.fg_bin_ranked_workspace_bytes(C, N, tile_w, tile_h) if BIN_RANKED else 0
        with _stage("bin_count"):
            diff = torch.empty(C * (tile_h + 1) * (tile_w + 1), **i32)
            if rank_bytes:
                wsr = _ws.get("bin_ranked", rank_bytes, dev)
                check(L.fg_bin_count_cells(C, N, ptr(order), ptr(means2d), ptr(radii), tile_size, tile_w, tile_h, ptr(diff),
                                           ptr(wsr), wsr.numel(), st))
            else:
                coarse_cnt = torch.empty(total, **i32)
                check(L.fg_bin_count(C, N, ptr(order), ptr(means2d), ptr(radii), tile_size, tile_w, tile_h, ptr(diff),
                                     ptr(coarse_cnt), st))
            ws2 = _ws.get("tile_scan", L.fg_bin_tile_scan_workspace_bytes(C, tile_w, tile_h), dev)
            check(L.fg_bin_tile_scan(C, tile_w, tile_h, ptr(diff), ptr(isect_offsets), n2[0:1].data_ptr(), ptr(ws2),
                                     ws2.numel(), st))
            if rank_bytes:
                coarse_offsets = torch.empty(C * cw * chh + 1, **i32)
                check(L.fg_bin_cell_scan(C, N, tile_w, tile_h, ptr(n_vis_dev), ptr(wsr), wsr.numel(), ptr(coarse_offsets),
                                         n2[1:2].data_ptr(), st))
            else:
                check(L.fg_exclusive_scan_i32(total, ptr(coarse_cnt), ptr(offsets), n2[1:2].data_ptr(), ptr(ws),
                                              ws.numel(), st))
        M, Mc = (int(v) for v in n2.tolist())  # the one host sync: sizes the list buffers
        assert M < 2**31, "too many tile intersections"
        fl = torch.empty(M, **i32)
        if M > 0:
            with _stage("coarse_sort"):
                cv = torch.empty(Mc, **i32)
                if rank_bytes:
                    check(L.fg_bin_ranked_emit(C, N, ptr(order), ptr(means2d), ptr(radii), tile_size, tile_w, tile_h,
                                               ptr(wsr), wsr.numel(), ptr(coarse_offsets), ptr(cv), st))
                else:
                    ck = torch.empty(Mc, **i32)
                    check(L.fg_bin_coarse_emit(C, N, ptr(order), ptr(means2d), ptr(radii), ptr(offsets), tile_size,
                                               tile_w, tile_h, ptr(ck), ptr(cv), st))
                    bits = max(1, int(math.ceil(math.log2(C * cw * chh))))
                    ck, cv = _sort_pairs(L, Mc, ck, cv, torch.empty_like(ck), torch.empty_like(cv), bits, dev, st, False)
                    coarse_offsets = torch.empty(C * cw * chh, **i32)
                    check(L.fg_isect_offsets_tiles(Mc, ptr(ck), C, cw, chh, ptr(coarse_offsets), st))
            with _stage("fine_bin"):
                check(L.fg_bin_fine(C, N, Mc, ptr(coarse_offsets), ptr(cv), ptr(means2d), ptr(radii), tile_size,
                                    tile_w, tile_h, ptr(isect_offsets), ptr(fl), st))
        return None, fl, isect_offsets, None

    # ---- two-level
    with _stage("scan"):
        cnt_sorted = torch.empty(total, **i32)
        check(L.fg_gather_i32(total, ptr(tiles_per_gauss), ptr(order), ptr(cnt_sorted), st))
        check(L.fg_exclusive_scan_i32(total, ptr(cnt_sorted), ptr(offsets), ptr(n_dev), ptr(ws), ws.numel(), st))
    M = int(n_dev.item())  # host sync: sizes the intersection buffers
    assert M < 2**31, "too many tile intersections"
    tk = torch.empty(M, **i32)
    fl = torch.empty(M, **i32)
    if M > 0:
        with _stage("emit"):
            check(L.fg_isect_emit_tiles(C, N, ptr(order), ptr(means2d), ptr(radii), ptr(offsets), tile_size, tile_w,
                                        tile_h, ptr(tk), ptr(fl), st))
        bits = max(1, int(math.ceil(math.log2(C * tile_w * tile_h))))
        with _stage("sort"):
            tk, fl = _sort_pairs(L, M, tk, fl, torch.empty_like(tk), torch.empty_like(fl), bits, dev, st, False)
    with _stage("offsets"):
        check(L.fg_isect_offsets_tiles(M, ptr(tk), C, tile_w, tile_h, ptr(isect_offsets), st))
    return None, fl, isect_offsets, tk


@torch.no_grad()
def isect_ids_from_tiles(tile_keys: Tensor, flatten_ids: Tensor, depths: Tensor, tile_w: int, tile_h: int) -> Tensor:
    """The reference's sorted 64-bit keys (gsplat ``meta["isect_ids"]``) rebuilt from the two-level result."""
    L = _lib.lib()
    M = flatten_ids.shape[0]
    out = torch.empty(M, dtype=torch.int64, device=flatten_ids.device)
    check(L.fg_isect_ids_from_tiles(M, ptr(tile_keys), ptr(flatten_ids), ptr(depths.contiguous()), tile_w, tile_h,
                                    ptr(out), _stream()))
    return out


class _Meta(dict):
    """``meta`` dict whose rarely-read entries (``isect_ids``) are materialised on first access."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._lazy = {}

    def lazy(self, key, fn):
        self._lazy[key] = fn

    def __getitem__(self, key):
        if not dict.__contains__(self, key) and key in self._lazy:
            dict.__setitem__(self, key, self._lazy.pop(key)())
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def get(self, key, default=None):
        return self[key] if key in self else default


# --------------------------------------------------------------------------- compositing
class _Rasterize(torch.autograd.Function):
    """fg_rasterize_fwd / fg_rasterize_bwd over <= 8 channels.

    Channels [0,split) come back as ``render``, channels [split,CH) as ``render2`` (the flow image);
    ``ed_channel`` is normalised by alpha inside the kernel.  ``opacities`` is ``[N]`` (shared by all
    cameras) or ``[C,N]``."""

    @staticmethod
    def forward(ctx, means2d, conics, feat, opacities, backgrounds, isect_offsets, flatten_ids, width, height,
                tile_size, absgrad, split, ed_channel, chunked, flow_affine=None):
        L = _lib.lib()
        ctx.chunked = chunked
        C = isect_offsets.shape[0]
        CH = feat.shape[-1]
        NN = feat.numel() // CH  # C*N (or nnz when packed)
        dev = feat.device
        means2d_c, conics_c, feat_c, opac_c = (t.contiguous() for t in (means2d, conics, feat, opacities))
        opac_shared = int(opac_c.numel() != NN)
        n_shared = opac_c.numel()
        bg = None if backgrounds is None else backgrounds.contiguous()
        aff = None if flow_affine is None else flow_affine.contiguous()
        render = torch.empty(C, height, width, split, device=dev)
        render2 = torch.empty(C, height, width, CH - split, device=dev) if split < CH else None
        alphas = torch.empty(C, height, width, 1, device=dev)
        last_ids = torch.empty(C, height, width, dtype=torch.int32, device=dev)
        M = flatten_ids.shape[0]
        n_arg = n_shared if opac_shared else NN
        with _stage("rasterize_fwd"):
            check(L.fg_rasterize_fwd(C, n_arg, CH, width, height, tile_size, ptr(means2d_c), ptr(conics_c),
                                     ptr(feat_c), ptr(opac_c), ptr(bg), ptr(aff), split, split, ed_channel,
                                     opac_shared, ptr(isect_offsets), ptr(flatten_ids), M, ptr(render),
                                     ptr(render2), ptr(alphas), ptr(last_ids), _stream()))
        ctx.save_for_backward(means2d_c, conics_c, feat_c, opac_c, bg, isect_offsets, flatten_ids, alphas, last_ids,
                              render if ed_channel >= 0 else None, aff)
        ctx.dims = (C, NN, CH, width, height, tile_size, absgrad, split, ed_channel, opac_shared, n_shared)
        ctx.means2d_obj = means2d  # the very tensor the caller holds as meta["means2d"] (model.py:869-871)
        ctx.mark_non_differentiable(last_ids)
        if render2 is None:
            render2 = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(render2)
        return render, render2, alphas, last_ids

    @staticmethod
    def backward(ctx, v_render, v_render2, v_alphas, _v_last):
        L = _lib.lib()
        (means2d, conics, feat, opac, bg, isect_offsets, flatten_ids, alphas, last_ids, render,
         aff) = ctx.saved_tensors
        C, NN, CH, width, height, tile_size, absgrad, split, ed_channel, opac_shared, n_shared = ctx.dims
        dev = feat.device

        def c(t):
            return None if t is None else t.contiguous()

        v_render, v_alphas = c(v_render), c(v_alphas)
        v_render2 = c(v_render2) if split < CH else None
        # one zero-filled arena for the five atomically accumulated gradient tensors
        sizes = [means2d.numel(), means2d.numel() if absgrad else 0, conics.numel(), feat.numel(), opac.numel(),
                 0 if aff is None else aff.numel()]
        arena = torch.zeros(sum(sizes), device=dev)
        parts = torch.split(arena, sizes)
        v_aff = None if aff is None else parts[5].view_as(aff)
        v_means2d = parts[0].view_as(means2d)
        v_abs = parts[1].view_as(means2d) if absgrad else None
        v_conics, v_feat, v_opac = parts[2].view_as(conics), parts[3].view_as(feat), parts[4].view_as(opac)
        M = flatten_ids.shape[0]
        with _stage("rasterize_bwd"):
            check(L.fg_rasterize_bwd(C, n_shared if opac_shared else NN, CH, width, height, tile_size, ptr(means2d),
                                     ptr(conics), ptr(feat), ptr(opac), ptr(bg), ptr(aff), split, split, ed_channel,
                                     opac_shared, ptr(isect_offsets), ptr(flatten_ids), M, ptr(render), ptr(alphas),
                                     ptr(last_ids), ptr(v_render), ptr(v_render2), ptr(v_alphas), ptr(v_means2d),
                                     ptr(v_abs), ptr(v_conics), ptr(v_feat), ptr(v_opac), ptr(v_aff), _stream()))
        if absgrad:
            obj = ctx.means2d_obj
            prev = getattr(obj, "absgrad", None) if ctx.chunked else None
            obj.absgrad = v_abs if prev is None else prev + v_abs
        v_bg = None
        if bg is not None and ctx.needs_input_grad[4]:
            vr = v_render if v_render is not None else torch.zeros(C, height, width, split, device=dev)
            if split < CH:
                vr2 = v_render2 if v_render2 is not None else torch.zeros(C, height, width, CH - split, device=dev)
                vr = torch.cat([vr, vr2], -1)
            if ed_channel >= 0:  # the ED channel's background is zero by construction
                vr = vr.clone()
                vr[..., ed_channel] = 0
            v_bg = (vr * (1.0 - alphas)).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_feat, v_opac, v_bg) + (None,) * 9 + (v_aff,)


def rasterize_to_pixels(means2d, conics, colors, opacities, image_width, image_height, tile_size, isect_offsets,
                        flatten_ids, backgrounds=None, absgrad=False, return_last_ids=False):
    """Composite ``colors [.., D]`` (any D; chunks of 8 channels per pass).  gsplat-compatible helper."""
    D = colors.shape[-1]
    outs, alphas, last_ids = [], None, None
    for s in range(0, D, MAX_CH):
        e = min(D, s + MAX_CH)
        bg = None if backgrounds is None else backgrounds[..., s:e]
        r, _, a, last_ids = _Rasterize.apply(means2d, conics, colors[..., s:e], opacities, bg, isect_offsets,
                                             flatten_ids, image_width, image_height, tile_size, absgrad, e - s, -1,
                                             D > MAX_CH)
        outs.append(r)
        alphas = a if alphas is None else alphas
    render = outs[0] if len(outs) == 1 else torch.cat(outs, -1)
    return (render, alphas, last_ids) if return_last_ids else (render, alphas)


# --------------------------------------------------------------------------- boundary
@_lib.on_device_of("means")
def rasterization(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [(C,) N, D] or [N, K, 3]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
    # ---- rendered Gaussian flow (north_star extension)
    means_next: Optional[Tensor] = None,
    quats_next: Optional[Tensor] = None,
    scales_next: Optional[Tensor] = None,
    flow_mode: str = "mean",
) -> Tuple[Tensor, Tensor, Dict]:
    """Render N Gaussians into C cameras; returns ``(render [C,H,W,X], alpha [C,H,W,1], meta)``.

    X = 3 ("RGB"), 4 ("RGB+D"/"RGB+ED", depth last) or 1 ("D"/"ED"); with ``sh_degree=None``
    the D colour channels are rendered as given.  ``meta["means2d"]`` is a graph tensor
    (``retain_grad()`` works) that carries ``.absgrad`` after backward when ``absgrad=True``.
    """
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in ["RGB", "D", "ED", "RGB+D", "RGB+ED"], render_mode
    assert rasterize_mode in ["classic", "antialiased"], rasterize_mode
    assert flow_mode in ["mean", "cov"], flow_mode
    assert tile_size == 16, "tile_size must be 16 (freegaussian_model.py:806)"
    if sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (
            colors.dim() == 3 and colors.shape[:2] == (C, N)
        ), colors.shape
    else:
        assert colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3, colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[1], colors.shape
        assert 0 <= sh_degree <= 3, "sh_degree must be in 0..3"
    if backgrounds is not None:
        assert backgrounds.shape[0] == C, backgrounds.shape
    if covars is not None or distributed or camera_model != "pinhole" or sparse_grad:
        raise NotImplementedError(
            "covars / distributed / non-pinhole cameras / sparse_grad are never used by the reference "
            "(freegaussian_model.py:847-868) and are not implemented"
        )
    if viewmats.requires_grad or Ks.requires_grad:
        raise NotImplementedError("camera gradients: the reference's camera optimizer is 'off' (model.py:120)")
    _require_cuda(means=means, quats=quats, scales=scales, opacities=opacities, colors=colors, viewmats=viewmats,
                  Ks=Ks, backgrounds=backgrounds, means_next=means_next)
    if means_next is not None:
        assert means_next.shape == (N, 3), means_next.shape

    want_depth = render_mode in ("D", "ED", "RGB+D", "RGB+ED")
    only_depth = render_mode in ("D", "ED")
    cfg = dict(width=int(width), height=int(height), eps2d=float(eps2d), near_plane=float(near_plane),
               far_plane=float(far_plane), radius_clip=float(radius_clip), tile_size=int(tile_size),
               sh_degree=None if only_depth else sh_degree, want_depth=want_depth,
               antialiased=rasterize_mode == "antialiased", flow_cov=flow_mode == "cov" and means_next is not None)
    n_feat = (0 if only_depth else (3 if sh_degree is not None else colors.shape[-1])) + int(want_depth) + (
        2 if means_next is not None else 0)
    cfg["fused"] = bool(FUSED_CALLS and SORT_MODE == "binned" and not stage_timer.enabled and not packed
                        and n_feat <= MAX_CH)
    proj_colors = None if only_depth else colors
    radii, means2d, depths, conics, comps, feat, tiles, flow_affine, opacities = _Project.apply(
        means, quats, scales, proj_colors, means_next, quats_next, scales_next, viewmats, Ks, cfg, opacities)
    if not cfg["flow_cov"]:
        flow_affine = None
    n_user = feat.shape[-1] - (2 if means_next is not None else 0)

    # classic: one opacity per Gaussian shared by all cameras ([N], no expand/copy); antialiased: per (c,n)
    opac = opacities * comps if rasterize_mode == "antialiased" else opacities

    if backgrounds is not None:
        if only_depth:
            backgrounds = torch.zeros(C, 1, device=means.device)
        elif want_depth:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, device=means.device)], -1)
        if means_next is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 2, device=means.device)], -1)

    tile_w = math.ceil(width / tile_size)
    tile_h = math.ceil(height / tile_size)
    front = cfg.pop("_front", None)
    if front is not None:
        isect_ids, tile_keys = None, None
        isect_offsets = front["isect_offsets"]
        flatten_ids = front["flatten_ids"]  # built inside the projection call, right after its host sync
    else:
        isect_ids, flatten_ids, isect_offsets, tile_keys = isect_tiles(means2d, radii, depths, tiles, tile_size, tile_w,
                                                                       tile_h, mode=SORT_MODE)

    meta = _Meta()
    if isect_ids is None:
        flat_unpacked, depths_unpacked = flatten_ids, depths.detach()

        def _rebuild_ids():
            tk = tile_keys
            if tk is None:  # binned mode: the tile of every list entry follows from the offsets
                o = isect_offsets.reshape(-1).long()
                counts = torch.diff(o, append=o.new_tensor([flat_unpacked.numel()]))
                tk = torch.repeat_interleave(torch.arange(o.numel(), device=o.device), counts).to(torch.int32)
            return isect_ids_from_tiles(tk, flat_unpacked, depths_unpacked, tile_w, tile_h)

        meta.lazy("isect_ids", _rebuild_ids)
    else:
        meta["isect_ids"] = isect_ids
    if packed:
        # compact per-Gaussian tensors to the visible (c,n) pairs in ascending order (Appendix A.8)
        needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (means2d, depths, conics, feat, opac))
        if not needs_grad and C * N > 0:
            # the preprocess callers (no gradients): three kernels -- flags + scan, one gather of every tensor, list remap
            L = _lib.lib()
            dev, st = means.device, _stream()
            total = C * N
            offs = torch.empty(total, dtype=torch.int32, device=dev)
            nnz_dev = torch.empty(1, dtype=torch.int64, device=dev)
            ws = _ws.get("pack", L.fg_pack_workspace_bytes(total), dev)
            check(L.fg_pack_plan(total, ptr(radii), ptr(offs), ptr(nnz_dev), ptr(ws), ws.numel(), st))
            nnz = int(nnz_dev.item())  # sizes the packed tensors (gsplat's packed projection has the same host read)
            CHf = feat.shape[-1]
            opac_c = opac.contiguous()
            new = lambda *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device=dev)  # noqa: E731
            radii_p, means2d_p, depths_p, conics_p = new(nnz, dtype=torch.int32), new(nnz, 2), new(nnz), new(nnz, 3)
            feat_p, opac_p = new(nnz, CHf), new(nnz)
            aff_p = new(nnz, 4) if flow_affine is not None else None
            cam_ids, gauss_ids = new(nnz, dtype=torch.int64), new(nnz, dtype=torch.int64)
            check(L.fg_pack_gather(C, N, CHf, ptr(radii), ptr(offs), ptr(means2d.detach().contiguous()),
                                   ptr(depths.detach().contiguous()), ptr(conics.detach().contiguous()),
                                   ptr(feat.detach().contiguous()), ptr(opac_c.detach()), int(opac_c.dim() == 1),
                                   ptr(flow_affine), ptr(radii_p), ptr(means2d_p), ptr(depths_p), ptr(conics_p), ptr(feat_p),
                                   ptr(opac_p), ptr(aff_p), ptr(cam_ids), ptr(gauss_ids), st))
            flatten_ids = flatten_ids.clone()
            check(L.fg_pack_remap(flatten_ids.numel(), ptr(offs), ptr(flatten_ids), st))
            radii, means2d, depths, conics, feat, opac, flow_affine = radii_p, means2d_p, depths_p, conics_p, feat_p, opac_p, aff_p
            meta["camera_ids"], meta["gaussian_ids"] = cam_ids, gauss_ids
        else:  # differentiable form (torch indexing): gradients flow back to the unpacked projection outputs
            vis = (radii > 0).reshape(-1)
            idx = torch.nonzero(vis).squeeze(-1)
            remap = (torch.cumsum(vis, 0, dtype=torch.int32) - 1).to(torch.int32)
            flatten_ids = remap[flatten_ids.long()].contiguous()
            means2d = means2d.reshape(C * N, 2)[idx]
            depths = depths.reshape(C * N)[idx]
            conics = conics.reshape(C * N, 3)[idx]
            feat = feat.reshape(C * N, -1)[idx]
            opac = (opac if opac.dim() == 2 else opac[None].expand(C, N)).reshape(C * N)[idx]
            if flow_affine is not None:
                flow_affine = flow_affine.reshape(C * N, 4)[idx]
            radii = radii.reshape(C * N)[idx]
            meta["camera_ids"] = idx // N
            meta["gaussian_ids"] = idx % N

    CH = feat.shape[-1]
    ed_channel = n_user - 1 if render_mode in ("ED", "RGB+ED") else -1
    flow = None
    if CH <= MAX_CH:
        render, flow_img, alphas, last_ids = _Rasterize.apply(
            means2d, conics, feat, opac, backgrounds, isect_offsets, flatten_ids, width, height, tile_size, absgrad,
            n_user, ed_channel, False, flow_affine)
        if means_next is not None:
            flow = flow_img
    else:  # many user colour channels: chunks of 8, normalisation / split done by torch
        if flow_affine is not None:
            raise NotImplementedError("flow_mode='cov' with more than 8 composited channels")
        opac_full = opac if opac.dim() == 2 or packed else opac[None].expand(C, N)
        render_all, alphas, last_ids = rasterize_to_pixels(means2d, conics, feat, opac_full, width, height, tile_size,
                                                           isect_offsets, flatten_ids, backgrounds=backgrounds,
                                                           absgrad=absgrad, return_last_ids=True)
        render = render_all[..., :n_user]
        if ed_channel >= 0:
            render = torch.cat([render[..., :-1], render[..., -1:] / alphas.clamp(min=1e-10)], -1)
        if means_next is not None:
            flow = render_all[..., n_user:]

    meta.update({
        "radii": radii, "means2d": means2d, "depths": depths, "conics": conics, "opacities": opac,
        "tile_width": tile_w, "tile_height": tile_h, "tiles_per_gauss": tiles,
        "flatten_ids": flatten_ids, "isect_offsets": isect_offsets, "width": width, "height": height,
        "tile_size": tile_size, "n_cameras": C, "last_ids": last_ids,
    })
    if means_next is not None:
        meta["flow"] = flow
    return render, alphas, meta
