"""Deformation and control networks of FreeGaussian on the B200 tensor cores.

``DeformNetwork`` mirrors ``FreeGaussianDeformableModel`` (``freegaussian/freegaussian_model.py:1054-1114``): same
constructor arguments, same parameter names and shapes (a reference ``state_dict`` loads unchanged), same
``forward(x, t) -> (d_xyz [N,4,4], rotation [N,4], scaling [N,3])``.  ``deform_gaussians`` additionally fuses the
lines that consume those outputs (``freegaussian_model.py:836-845``) and returns the ``means, scales, quats`` handed
to ``rasterization``.  ``ControlNetwork`` mirrors the stage-2 ``FreeGaussianControllableModel`` (``:1117-1145``) the same way.

Every ``nn.Linear`` of the trunk runs on tcgen05 / TMEM / TMA (csrc/mlp.cu) in error-compensated 3xTF32 (fp32-accurate:
the reference computes these layers in fp32): forward and data gradient in ``fg_mlp_linear``, weight and bias gradients
in ``fg_mlp_wgrad`` (split-K, MN-major operands), the heads' included.
There is no CPU path.

The reference always evaluates the network with one time value per call (``camera.times.expand(N, -1)``,
``freegaussian_model.py:836``), so the time branch (embedding + ``timenet``) is evaluated once on a single row with
ordinary torch ops and broadcast inside the embedding kernel; its gradient comes back through the bias gradients
of the two layers that read it.
"""

from __future__ import annotations

from typing import List, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import MLP_EMBED_LD, MLP_HEAD_LD, MlpPackSegment, check, ptr

_W = 256
_D = 8
_SKIP = _D // 2  # the embedding is concatenated again after this layer (freegaussian_model.py:1058, 1100-1101)
_HEADS = (("branch_w", 3), ("branch_v", 3), ("gaussian_rotation", 4), ("gaussian_scaling", 3))
# backward over the rows with a non-zero incoming gradient only, when they are fewer than this fraction of all rows
SPARSE_BACKWARD = True
SPARSE_BACKWARD_MAX_FRACTION = 0.75
_CAP_SLACK, _CAP_EXTRA = 1.12, 1024  # head-room of the guessed capacity over the previous count
_ACTIVE_ROWS: dict = {}   # (device, rows, embedding width) -> active rows of the previous backward of that shape
_PINNED: dict = {}
STATS = {"sparse_backward": 0, "dense_backward": 0, "overflow_redo": 0}  # counters (bench.py reports them)


def _pinned_slot(key) -> Tensor:
    if key not in _PINNED:
        _PINNED[key] = torch.zeros(1, dtype=torch.int64).pin_memory()
    return _PINNED[key]


def _embed(x: Tensor, multires: int) -> Tensor:
    """Embedder.embed (utils.py:27-56) with torch ops; used for the single time row only."""
    out = [x]
    for k in range(multires):
        out += [torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))]
    return torch.cat(out, -1)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _linear(mode, M, n_out, a0, k0, a1, k1, w, bias, mask_in, out, mask_out):
    """``w`` is the (hi, lo) pair of a packed weight; activations are plain fp32."""
    check(_lib.lib().fg_mlp_linear(mode, M, n_out, ptr(a0), k0, ptr(a1), k1, ptr(w[0]), ptr(w[1]), ptr(bias), ptr(mask_in),
                                   ptr(out), ptr(mask_out), _stream()))


class _Spec:
    """Shape of one network: embedding row length (96 / 128), used embedding channels, heads, whether the data gradient
    has to reach the embedded points (stage 2 does not detach them)."""

    def __init__(self, emb_ld: int, emb_ch: int, multires: int, heads, input_grad: bool):
        self.emb_ld, self.emb_ch, self.multires, self.heads, self.input_grad = emb_ld, emb_ch, multires, tuple(heads), input_grad
        assert emb_ch <= emb_ld and emb_ld % 32 == 0 and sum(o for _, o in heads) <= MLP_HEAD_LD
        assert not input_grad or emb_ld == 128, "the embedding-gradient GEMM is built for 128 columns"


def _pack_plan(spec: _Spec):
    """The weight-packing table as plain data: one tuple per segment,
    ``(param index, source column 0, columns, destination, destination column 0, transpose, destination row 0)`` with
    destination one of ``("w", layer) | ("wt", layer) | ("w_head",) | ("wt_head",) | ("wt_emb",)``.
    Pure host logic (tests/test_deform.py::test_pack_plan_rebuilds_the_layers applies it with numpy)."""
    emb_ch = spec.emb_ch
    plan = []
    for i in range(_D):
        p = 2 * i
        if i == 0:
            plan.append((p, 0, emb_ch, ("w", 0), 0, False, 0))
            if spec.input_grad:
                plan.append((p, 0, emb_ch, ("wt_emb",), 0, True, 0))
        elif i == _SKIP + 1:  # reference input order [embedding | h]; operand order [h | embedding]
            plan.append((p, emb_ch, _W, ("w", i), 0, False, 0))
            plan.append((p, 0, emb_ch, ("w", i), _W, False, 0))
            plan.append((p, emb_ch, _W, ("wt", i), 0, True, 0))
            if spec.input_grad:
                plan.append((p, 0, emb_ch, ("wt_emb",), _W, True, 0))
        else:
            plan.append((p, 0, _W, ("w", i), 0, False, 0))
            plan.append((p, 0, _W, ("wt", i), 0, True, 0))
    row = 0
    for j, (_, o) in enumerate(spec.heads):
        p = 2 * _D + 2 * j
        plan.append((p, 0, _W, ("w_head",), 0, False, row))
        plan.append((p, 0, _W, ("wt_head",), row, True, 0))
        row += o
    return plan


def _packed_shapes(spec: _Spec):
    """Shapes of the operand buffers the plan writes into."""
    ld = spec.emb_ld
    shapes = {("w", i): (_W, ld if i == 0 else (_W + ld if i == _SKIP + 1 else _W)) for i in range(_D)}
    shapes.update({("wt", i): (_W, _W) for i in range(1, _D)})
    shapes[("w_head",)] = (MLP_HEAD_LD, _W)
    shapes[("wt_head",)] = (_W, MLP_HEAD_LD)
    if spec.input_grad:
        shapes[("wt_emb",)] = (ld, 2 * _W)  # wt_emb[e, :] = [W_0[:, e] | W_skip[:, e]]
    return shapes


class _Packed:
    """Operand buffers of one parameter set: padded, reordered, hi/lo split, plus the transposes for the data gradient."""

    def __init__(self, params: List[Tensor], spec: _Spec):
        dev = params[0].device
        shapes = _packed_shapes(spec)
        total = sum(r * c for r, c in shapes.values())
        arena = torch.zeros(2, total, device=dev, dtype=torch.float32)  # one memset: the pads stay zero
        bufs, o = {}, 0
        for k, (r, c) in shapes.items():  # every size is a multiple of 32 floats, so the slices stay 128-byte aligned
            bufs[k] = (arena[0, o:o + r * c].view(r, c), arena[1, o:o + r * c].view(r, c))
            o += r * c
        plan = _pack_plan(spec)
        assert len(plan) <= _lib.MLP_PACK_MAX_SEGMENTS
        arr = (MlpPackSegment * len(plan))()
        for s, (p, col0, cols, dst, dst_col0, transpose, row0) in zip(arr, plan):
            src = params[p]
            assert src.is_contiguous()
            hi, lo = bufs[dst]
            s.src = src.data_ptr()
            s.dst_hi = hi.data_ptr() + row0 * hi.shape[1] * 4
            s.dst_lo = lo.data_ptr() + row0 * lo.shape[1] * 4
            s.src_ld, s.src_col0, s.rows, s.cols = src.shape[1], col0, src.shape[0], cols
            s.dst_ld, s.dst_col0, s.transpose = hi.shape[1], dst_col0, int(transpose)
        check(_lib.lib().fg_mlp_pack(len(plan), arr, _stream()))
        self.w = [bufs[("w", i)] for i in range(_D)]
        self.wt = [None] + [bufs[("wt", i)] for i in range(1, _D)]  # wt[l][i, o] = W_l[o, i] over the hidden inputs
        self.w_head, self.wt_head = bufs[("w_head",)], bufs[("wt_head",)]
        self.wt_emb = bufs.get(("wt_emb",))
        self.bias = [b.contiguous() for b in params[1:2 * _D:2]]
        hb = torch.cat([params[2 * _D + 2 * j + 1] for j in range(len(spec.heads))])
        self.bias_head = torch.cat([hb, hb.new_zeros(MLP_HEAD_LD - hb.numel())])


class _ActivationPool:
    """Persistent activation workspaces of the trunk, keyed by (device, rows, embedding width): the embedding [N, ld], the
    eight hidden activations [8, N, 256] and their ReLU bit masks [8, N, 8] -- 9.6 GB at 1 M rows.  A forward pass leases
    one (allocating only when none is free), the lease goes back to the pool when the autograd context that saved it
    dies (after the backward, or when an inference result is dropped).  Without the pool every training iteration
    allocated and freed these tensors through the caching allocator, whose occasional cudaMalloc / cudaFree of
    gigabyte blocks showed up as 30-190 ms iterations (bench.py train_iter: `cuda_mallocs_in_timed_region`)."""

    def __init__(self):
        self.free = {}

    def lease(self, dev, N: int, ld: int):
        key = (str(dev), N, ld)
        stack = self.free.setdefault(key, [])
        if stack:
            bufs = stack.pop()
        else:
            bufs = (torch.empty(N, ld, device=dev, dtype=torch.float32),
                    torch.empty(_D, N, _W, device=dev, dtype=torch.float32),
                    torch.empty(_D, N, _W // 32, device=dev, dtype=torch.int32))
        return _Lease(self, key, bufs)

    def clear(self):
        self.free.clear()


class _Lease:
    def __init__(self, pool, key, bufs):
        self.pool, self.key, self.bufs = pool, key, bufs

    def __del__(self):
        try:
            stack = self.pool.free.setdefault(self.key, [])
            if len(stack) < 2:  # at most two idle workspaces per shape are kept
                stack.append(self.bufs)
        except Exception:  # interpreter shutdown
            pass


_POOL = _ActivationPool()
USE_ACTIVATION_POOL = True


class _Trunk(torch.autograd.Function):
    """x [N,3], x2 [N,3] or None (embedded like x, no gradient), t_emb [t_ch] or None, spec, parameters -> head [N, 32]."""

    @staticmethod
    def forward(ctx, x, x2, t_emb, spec, *params):
        L = _lib.lib()
        N = x.shape[0]
        dev = x.device
        ld = spec.emb_ld
        t_ch = t_emb.numel() if t_emb is not None else 0
        pk = _Packed(list(params), spec)
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        if USE_ACTIVATION_POOL and dev.type == "cuda" and N > 0:
            lease = _POOL.lease(dev, N, ld)
            e, h_all, masks = lease.bufs
        else:
            lease = None
            e, h_all = new(N, ld), new(_D, N, _W)
            masks = torch.empty(_D, N, _W // 32, device=dev, dtype=torch.int32)  # ReLU bit masks of all layers, one tensor
        check(L.fg_deform_embed(N, ptr(x), ptr(x2), ptr(t_emb), t_ch, spec.multires, ld, ptr(e), _stream()))
        hs: List[Tensor] = []
        prev = None
        for i in range(_D):
            out, bits = h_all[i], masks[i]
            if i == 0:
                _linear(_lib.MLP_RELU, N, _W, e, ld, None, 0, pk.w[0], pk.bias[0], None, out, bits)
            elif i == _SKIP + 1:
                _linear(_lib.MLP_RELU, N, _W, prev, _W, e, ld, pk.w[i], pk.bias[i], None, out, bits)
            else:
                _linear(_lib.MLP_RELU, N, _W, prev, _W, None, 0, pk.w[i], pk.bias[i], None, out, bits)
            prev = out
            hs.append(out)
        head = new(N, MLP_HEAD_LD)
        _linear(_lib.MLP_LINEAR, N, MLP_HEAD_LD, prev, _W, None, 0, pk.w_head, pk.bias_head, None, head, None)
        ctx.save_for_backward(x, e, masks, *hs, *params)
        ctx.pk, ctx.spec, ctx.t_ch = pk, spec, t_ch
        ctx.lease = lease  # the workspace returns to the pool when this context dies
        return head

    @staticmethod
    def backward(ctx, g_head):
        spec = ctx.spec
        ld = spec.emb_ld
        N_all = ctx.saved_tensors[0].shape[0]
        dev = g_head.device
        g_head = g_head.contiguous()
        L = _lib.lib()
        # Rows whose incoming gradient is exactly zero (Gaussians that were culled or never reached a pixel in this
        # step's views) contribute exactly nothing to any gradient: run the backward on the other rows only.
        # Their indices are compacted on the device (fg_rows_active) and the COUNT stays there: the row buffers get a
        # capacity guessed from the previous backward of this shape, fg_rows_gather zero-fills the rows past the count
        # (zero rows add nothing to any product below), and the count is only read back AFTER everything has been
        # enqueued -- if it exceeded the capacity (never in steady state: the capacity carries 12 % head-room) the
        # backward is redone densely.  No host synchronisation sits in front of the tensor-core kernels any more.
        if SPARSE_BACKWARD and N_all > 0:
            key = (str(dev), N_all, ld)
            ws = torch.empty(int(L.fg_rows_workspace_bytes(N_all)), device=dev, dtype=torch.uint8)
            idx_buf = torch.empty(N_all, device=dev, dtype=torch.int32)
            count_dev = torch.empty(1, device=dev, dtype=torch.int64)
            check(L.fg_rows_active(N_all, ptr(g_head), g_head.shape[1], ptr(idx_buf), ptr(count_dev), ptr(ws), ws.numel(),
                                   _stream()))
            guess, prev_cap = _ACTIVE_ROWS.get(key, (None, 0))
            pending = None
            if guess is None:  # first backward of this shape: one host read, exact capacity
                cap = int(count_dev.item())
            else:
                cap = min(N_all, -(-int(guess * _CAP_SLACK + _CAP_EXTRA) // 128) * 128)
                # hysteresis: keep the previous capacity unless the need grows beyond it or falls below 70 % of it, so
                # that the row buffers have the SAME sizes step after step and the caching allocator reuses them (a
                # capacity that follows the count row by row made it go to cudaMalloc now and then: 30 ms outliers)
                if prev_cap and 0.7 * prev_cap <= cap <= prev_cap:
                    cap = prev_cap
                if dev.type == "cuda":
                    pinned = _pinned_slot(key)
                    pinned.copy_(count_dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    pending = (pinned, ev)
                else:
                    pending = (count_dev, None)
            if cap < SPARSE_BACKWARD_MAX_FRACTION * N_all:
                STATS["sparse_backward"] += 1
                grads_out = _Trunk._backward_rows(ctx, g_head, (idx_buf, count_dev), cap)
            else:
                STATS["dense_backward"] += 1
                grads_out = _Trunk._backward_rows(ctx, g_head, None, N_all)
                cap = N_all
            if pending is not None:
                if pending[1] is not None:
                    pending[1].synchronize()
                count = int(pending[0].item())
                if count > cap:  # the guess was too small: redo on every row (exact, just slower)
                    STATS["overflow_redo"] += 1
                    grads_out = _Trunk._backward_rows(ctx, g_head, None, N_all)
            else:
                count = cap
            _ACTIVE_ROWS[key] = (count, cap)
            return grads_out
        return _Trunk._backward_rows(ctx, g_head, None, N_all)

    @staticmethod
    def _backward_rows(ctx, g_head, active, N):
        """The backward over N rows: every row (``active is None``) or the ``N``-row capacity buffer of the active rows
        ``active = (idx, count_dev)`` (rows past the count are zero)."""
        spec, pk, t_ch = ctx.spec, ctx.pk, ctx.t_ch
        ld, emb_ch = spec.emb_ld, spec.emb_ch
        saved = ctx.saved_tensors
        x, e, masks, hs, params = saved[0], saved[1], saved[2], saved[3:3 + _D], saved[3 + _D:]
        N_all = x.shape[0]
        dev = g_head.device
        grads: List[Tensor] = [None] * len(params)
        L = _lib.lib()
        idx = active[0] if active is not None else None

        def rows(width, dtype=torch.float32):
            return torch.empty(N, width, device=dev, dtype=dtype)

        new = lambda n_, w_: rows(w_)  # noqa: E731  (every [N, w] temporary of the backward)
        if idx is None:
            sel = lambda t: t  # noqa: E731
            m_in = lambda i: masks[i]  # noqa: E731
        else:
            def sel(t):
                out = rows(t.shape[1], t.dtype)
                check(L.fg_rows_gather(N, ptr(idx), ptr(active[1]), ptr(t), t.shape[1] * t.element_size(), ptr(out), _stream()))
                return out

            m_in = lambda i: sel(masks[i])  # noqa: E731
            g_head = sel(g_head)
        e = sel(e)
        h_in = lambda i: sel(hs[i])  # noqa: E731  activations of layer i (input of layer i + 1)
        # head gradients: dW_head^T [256, 32] = h_last^T . g_head, one pass of the same kernel; biases = column sums of g_head
        dw_head_t = torch.zeros(_W, MLP_HEAD_LD, device=dev, dtype=torch.float32)
        h_last = h_in(_D - 1)
        check(L.fg_mlp_wgrad(N, ptr(h_last), ptr(g_head), MLP_HEAD_LD, ptr(dw_head_t), MLP_HEAD_LD, 0, None, _stream()))
        del h_last
        db_head = g_head.sum(0)
        row = 0
        for j, (_, o) in enumerate(spec.heads):
            grads[2 * _D + 2 * j] = dw_head_t[:, row:row + o].t()
            grads[2 * _D + 2 * j + 1] = db_head[row:row + o]
            row += o
        # trunk: one zero-filled arena for everything fg_mlp_wgrad adds into
        arena = torch.zeros(_D * (_W * _W + _W) + 2 * _W * ld, device=dev, dtype=torch.float32)
        dw_h = [arena[i * _W * _W:(i + 1) * _W * _W].view(_W, _W) for i in range(_D)]  # dw_h[0] unused
        db = [arena[_D * _W * _W + i * _W:_D * _W * _W + (i + 1) * _W] for i in range(_D)]
        off = _D * (_W * _W + _W)
        dw_e = [arena[off + k * _W * ld:off + (k + 1) * _W * ld].view(_W, ld) for k in range(2)]
        st = _stream()

        def wgrad(dz, a, k_in, dw, db_):
            check(L.fg_mlp_wgrad(N, ptr(dz), ptr(a), k_in, ptr(dw), k_in, 0, ptr(db_) if db_ is not None else None, st))

        dz = new(N, _W)
        _linear(_lib.MLP_DGRAD, N, _W, g_head, MLP_HEAD_LD, None, 0, pk.wt_head, None, m_in(_D - 1), dz, None)
        g_t = torch.zeros(t_ch, device=dev) if t_ch else None
        dz_skip = None
        for i in range(_D - 1, -1, -1):
            if i == 0:
                wgrad(dz, e, ld, dw_e[0], db[0])
                grads[0] = dw_e[0][:, :emb_ch]
            elif i == _SKIP + 1:  # reference column order: [embedding | h]
                wgrad(dz, h_in(i - 1), _W, dw_h[i], db[i])
                wgrad(dz, e, ld, dw_e[1], None)
                grads[2 * i] = torch.cat([dw_e[1][:, :emb_ch], dw_h[i]], 1)
                dz_skip = dz
            else:
                wgrad(dz, h_in(i - 1), _W, dw_h[i], db[i])
                grads[2 * i] = dw_h[i]
            grads[2 * i + 1] = db[i]
            if t_ch and (i == 0 or i == _SKIP + 1):
                # every row reads the same t_emb, so its gradient is (column sums of dz) . W[:, t columns]
                g_t = g_t + db[i] @ params[2 * i][:, emb_ch - t_ch:emb_ch]
            if i > 0:
                dz_prev = new(N, _W)
                _linear(_lib.MLP_DGRAD, N, _W, dz, _W, None, 0, pk.wt[i], None, m_in(i - 1), dz_prev, None)
                dz = dz_prev
        g_x = None
        if spec.input_grad and ctx.needs_input_grad[0]:
            # d(loss)/d(embedding) = dz_0 . W_0[:, emb] + dz_skip . W_skip[:, emb] as ONE product over K = 512, then the VJP
            # of the positional embedding of x (sin / cos derivatives)
            de = new(N, ld)
            _linear(_lib.MLP_LINEAR, N, ld, dz, _W, dz_skip, _W, pk.wt_emb, torch.zeros(ld, device=dev), None, de, None)
            dx = new(N, 3)
            x_rows = x if idx is None else x.index_select(0, idx[:N].long())  # 12-byte rows: not a 16-byte multiple
            check(L.fg_deform_embed_bwd(N, ptr(x_rows), ptr(de), spec.multires, ld, ptr(dx), st))
            # rows past the count are zero and point at row 0: adding them changes nothing
            g_x = dx if idx is None else torch.zeros(N_all, 3, device=dev).index_add_(0, idx[:N].long(), dx)
        return (g_x, None, g_t, None, *grads)


class _TimeBranch(torch.autograd.Function):
    """t [1] (one time value) [, timenet W1, b1, W2, b2] -> t_emb: positional embedding (utils.py:27-56) and, for the
    blender variant, ``timenet`` (freegaussian_model.py:1066-1071) in ONE launch forward and one backward
    (``fg_time_branch_fwd/bwd``) instead of ~60 single-row torch kernels per training iteration."""

    @staticmethod
    def forward(ctx, t0, multires, *net):
        L = _lib.lib()
        dev = t0.device
        in_ch = 1 + 2 * multires
        emb = torch.empty(in_ch, device=dev, dtype=torch.float32)
        if not net:
            check(L.fg_time_branch_fwd(ptr(t0), multires, in_ch, 0, 0, None, None, None, None, ptr(emb), None, None, _stream()))
            return emb
        w1, b1, w2, b2 = (p.contiguous() for p in net)
        hidden, out_ch = w1.shape[0], w2.shape[0]
        assert w1.shape == (hidden, in_ch) and w2.shape == (out_ch, hidden)
        h = torch.empty(hidden, device=dev, dtype=torch.float32)
        out = torch.empty(out_ch, device=dev, dtype=torch.float32)
        check(L.fg_time_branch_fwd(ptr(t0), multires, in_ch, hidden, out_ch, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(emb), ptr(h),
                                   ptr(out), _stream()))
        ctx.save_for_backward(emb, h, w2)
        ctx.dims = (in_ch, hidden, out_ch)
        return out

    @staticmethod
    def backward(ctx, g):
        emb, h, w2 = ctx.saved_tensors
        in_ch, hidden, out_ch = ctx.dims
        dev = g.device
        dw1 = torch.empty(hidden, in_ch, device=dev)
        db1 = torch.empty(hidden, device=dev)
        dw2 = torch.empty(out_ch, hidden, device=dev)
        db2 = torch.empty(out_ch, device=dev)
        check(_lib.lib().fg_time_branch_bwd(in_ch, hidden, out_ch, ptr(emb), ptr(h), ptr(w2), ptr(g.contiguous()), ptr(dw1),
                                            ptr(db1), ptr(dw2), ptr(db2), _stream()))
        return None, None, dw1, db1, dw2, db2


class _Apply(torch.autograd.Function):
    """head [N,32], means, scales_log, quats -> (means', scales', quats')   (freegaussian_model.py:841-845)."""

    @staticmethod
    def forward(ctx, head, means, scales_log, quats):
        N = means.shape[0]
        head, means, scales_log, quats = head.contiguous(), means.contiguous(), scales_log.contiguous(), quats.contiguous()
        mo, so, qo = torch.empty_like(means), torch.empty_like(scales_log), torch.empty_like(quats)
        check(_lib.lib().fg_deform_apply_fwd(N, ptr(head), ptr(means), ptr(scales_log), ptr(quats), ptr(mo), ptr(so), ptr(qo),
                                             _stream()))
        ctx.save_for_backward(head, means, scales_log, quats)
        return mo, so, qo

    @staticmethod
    def backward(ctx, g_m, g_s, g_q):
        head, means, scales_log, quats = ctx.saved_tensors
        N = means.shape[0]
        zero = lambda g, ref: torch.zeros_like(ref) if g is None else g.contiguous()  # noqa: E731
        g_m, g_s, g_q = zero(g_m, means), zero(g_s, scales_log), zero(g_q, quats)
        v_head = torch.empty_like(head)
        v_m, v_s, v_q = torch.empty_like(means), torch.empty_like(scales_log), torch.empty_like(quats)
        check(_lib.lib().fg_deform_apply_bwd(N, ptr(head), ptr(means), ptr(scales_log), ptr(quats), ptr(g_m), ptr(g_s), ptr(g_q),
                                             ptr(v_head), ptr(v_m), ptr(v_s), ptr(v_q), _stream()))
        return v_head, v_m, v_s, v_q


def _skew(w: Tensor) -> Tensor:
    z = torch.zeros_like(w[:, 0])
    return torch.stack([z, -w[:, 2], w[:, 1], w[:, 2], z, -w[:, 0], -w[:, 1], w[:, 0], z], -1).reshape(-1, 3, 3)


def exp_se3(S: Tensor, theta: Tensor) -> Tensor:
    """Same contract as ``freegaussian/utils.py:137-159``: screw axis [N,6], magnitude [N,1] -> [N,4,4]."""
    w, v = S[:, :3], S[:, 3:]
    Wm = _skew(w)
    W2 = torch.bmm(Wm, Wm)
    eye = torch.eye(3, device=S.device, dtype=S.dtype).expand_as(Wm)
    th = theta.view(-1, 1, 1)
    R = eye + torch.sin(th) * Wm + (1.0 - torch.cos(th)) * W2
    p = torch.bmm(th * eye + (1.0 - torch.cos(th)) * Wm + (th - torch.sin(th)) * W2, v.unsqueeze(-1))
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=S.device, dtype=S.dtype).expand(Wm.shape[0], 1, 4)
    return torch.cat([torch.cat([R, p], -1), bottom], 1)


class DeformNetwork(nn.Module):
    """Drop-in for ``FreeGaussianDeformableModel`` (freegaussian_model.py:1054-1114) with the trunk on tcgen05."""

    def __init__(self, D: int = 8, W: int = 256, multires: int = 10, is_blender: bool = False):
        super().__init__()
        assert D == _D and W == _W, "the tensor-core kernels are built for the reference's D=8, W=256"
        self.D, self.W, self.multires = D, W, multires
        self.t_multires = 6 if is_blender else 10
        self.skips = [D // 2]
        self.is_blender = is_blender
        time_input_ch = 1 + 2 * self.t_multires
        xyz_input_ch = 3 + 6 * multires
        if is_blender:
            self.time_out = 30
            self.timenet = nn.Sequential(nn.Linear(time_input_ch, 256), nn.ReLU(inplace=True), nn.Linear(256, self.time_out))
            t_ch = self.time_out
        else:
            t_ch = time_input_ch
        self.input_ch = xyz_input_ch + t_ch
        assert self.input_ch <= MLP_EMBED_LD
        self.linear = nn.ModuleList(
            [nn.Linear(self.input_ch, W)]
            + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.branch_w = nn.Linear(W, 3)
        self.branch_v = nn.Linear(W, 3)
        self.gaussian_rotation = nn.Linear(W, 4)
        self.gaussian_scaling = nn.Linear(W, 3)

    def _params(self) -> List[Tensor]:
        ps: List[Tensor] = []
        for lin in self.linear:
            ps += [lin.weight, lin.bias]
        for name, _ in _HEADS:
            lin = getattr(self, name)
            ps += [lin.weight, lin.bias]
        return ps

    def _time_row(self, t: Tensor) -> Tensor:
        """One time value per call (freegaussian_model.py:836 expands ``camera.times``); returns t_emb [t_ch]."""
        if t.numel() > 1 and not (t.dim() == 2 and t.stride(0) == 0):
            raise ValueError("DeformNetwork evaluates ONE time value per call, as the reference does with "
                             "`camera.times.expand(N, -1)`: pass a [1,1] tensor or an expanded view, not per-row times")
        if not t.is_cuda and not getattr(_lib.lib(), "host_memory_model", False):  # (tests/fake_mlp_lib.py runs on host memory)
            raise RuntimeError("DeformNetwork: `t` is not a CUDA tensor (there is no CPU path)")
        t0 = t.reshape(-1)[:1].to(torch.float32).contiguous()
        if self.is_blender:
            return _TimeBranch.apply(t0, self.t_multires, self.timenet[0].weight, self.timenet[0].bias,
                                     self.timenet[2].weight, self.timenet[2].bias)
        return _TimeBranch.apply(t0, self.t_multires)

    @_lib.on_device_of("x")
    def head(self, x: Tensor, t: Tensor) -> Tensor:
        """[N, 32]: branch_w (3) | branch_v (3) | gaussian_rotation (4) | gaussian_scaling (3) | zero padding."""
        if not x.is_cuda:
            raise RuntimeError("DeformNetwork: `x` is not a CUDA tensor (there is no CPU path)")
        assert x.ndim == 2 and x.shape[1] == 3 and x.dtype == torch.float32
        spec = _Spec(MLP_EMBED_LD, self.input_ch, self.multires, _HEADS, input_grad=False)
        return _Trunk.apply(x.detach().contiguous(), None, self._time_row(t).contiguous(), spec, *self._params())

    def forward(self, x: Tensor, t: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        h = self.head(x, t)
        w, v = h[:, 0:3], h[:, 3:6]
        theta = torch.norm(w, dim=-1, keepdim=True)
        w = w / theta + 1e-5
        v = v / theta + 1e-5
        d_xyz = exp_se3(torch.cat([w, v], -1), theta)
        return d_xyz, h[:, 6:10], h[:, 10:13]

    @_lib.on_device_of("means")
    def deform_gaussians(self, means: Tensor, scales_log: Tensor, quats: Tensor, t: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """freegaussian_model.py:836-845 in one call: returns the (means, scales, quats) passed to ``rasterization``."""
        return _Apply.apply(self.head(means, t), means, scales_log, quats)


_CONTROL_HEADS = (("d_xyz", 3), ("d_rot", 4), ("d_scale", 3))
_CONTROL_LD = 128


class ControlNetwork(nn.Module):
    """Drop-in for ``FreeGaussianControllableModel`` (freegaussian_model.py:1117-1145), the stage-2 network: the same trunk
    on the embedded control points and the embedded per-point control value.  ``x`` keeps its gradient (stage 2 passes
    ``means_crop[mask]`` without detaching, freegaussian_control_model.py:122, 143); ``value`` is built under ``no_grad``
    there (:127-142) and gets none."""

    def __init__(self, D: int = 8, W: int = 256, multires: int = 10):
        super().__init__()
        assert D == _D and W == _W, "the tensor-core kernels are built for the reference's D=8, W=256"
        self.D, self.W, self.multires = D, W, multires
        self.skips = [D // 2]
        self.input_ch = 2 * (3 + 6 * multires)
        assert self.input_ch <= _CONTROL_LD
        self.linear = nn.ModuleList(
            [nn.Linear(self.input_ch, W)]
            + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.d_xyz = nn.Linear(W, 3)
        self.d_scale = nn.Linear(W, 3)
        self.d_rot = nn.Linear(W, 4)

    def _params(self) -> List[Tensor]:
        ps: List[Tensor] = []
        for lin in self.linear:
            ps += [lin.weight, lin.bias]
        for name, _ in _CONTROL_HEADS:
            lin = getattr(self, name)
            ps += [lin.weight, lin.bias]
        return ps

    @_lib.on_device_of("x")
    def forward(self, x: Tensor, value: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """(d_xyz [N,3], d_rot [N,4], d_scale [N,3]) -- the reference's return order (:1144-1145)."""
        if not x.is_cuda:
            raise RuntimeError("ControlNetwork: `x` is not a CUDA tensor (there is no CPU path)")
        assert x.ndim == 2 and x.shape[1] == 3 and value.shape == x.shape and x.dtype == torch.float32
        spec = _Spec(_CONTROL_LD, self.input_ch, self.multires, _CONTROL_HEADS, input_grad=True)
        h = _Trunk.apply(x.contiguous(), value.detach().contiguous().to(torch.float32), None, spec, *self._params())
        return h[:, 0:3], h[:, 3:7], h[:, 7:10]
