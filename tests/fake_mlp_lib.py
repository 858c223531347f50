"""A host-memory model of the deformation-network entry points of include/fg_api.h, for CPU tests of the HOST logic.

TEST INFRASTRUCTURE ONLY.  `freegaussian_b200/deform.py` drives the network through raw pointers and the C ABI
(`fg_mlp_pack`, `fg_deform_embed`, `fg_mlp_linear`, `fg_mlp_wgrad`, `fg_deform_embed_bwd`, `fg_deform_apply_fwd/bwd`).
`FakeMlpLib` implements the documented semantics of those calls with numpy / torch on host pointers, so that the Python
orchestration (packing plan, skip-layer operand order, gradient assembly, sparse backward, time-gradient shortcut) can run
end to end on CPU tensors and be compared with the oracle.  It is an executable reading of the header's comments, nothing in
the product imports it, and the arithmetic is plain float64 -> float32 (the 3xTF32 scheme is the kernels' business).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

HEAD_LD = 32


def _arr(ptr, shape, dtype=np.float32):
    """numpy view of host memory at `ptr` (an int or c_void_p value)."""
    if ptr is None:
        return None
    n = int(np.prod(shape))
    ctype = {np.float32: ctypes.c_float, np.int32: ctypes.c_int32, np.uint32: ctypes.c_uint32}[dtype]
    buf = (ctype * n).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def _tf32_hi(x):
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return b.view(np.float32)


class FakeMlpLib:
    host_memory_model = True  # lets deform.py's "CUDA tensors only" guard of the time branch through

    def __init__(self):
        self.calls = []

    # ---- fg_mlp_pack
    def fg_mlp_pack(self, n, segs, stream):
        self.calls.append("pack")
        for i in range(n):
            s = segs[i]
            src = _arr(s.src, (s.rows, s.src_ld))[:, s.src_col0:s.src_col0 + s.cols]
            block = np.ascontiguousarray(src.T if s.transpose else src)
            hi = _tf32_hi(block)
            r, c = block.shape
            for dst_ptr, data in ((s.dst_hi, hi), (s.dst_lo, block - hi)):
                if not dst_ptr:
                    continue
                # dst_ptr may already sit at a row offset of its buffer: address the rows relative to it
                flat = _arr(dst_ptr, ((r - 1) * s.dst_ld + s.dst_col0 + c,))
                for rr in range(r):
                    flat[rr * s.dst_ld + s.dst_col0: rr * s.dst_ld + s.dst_col0 + c] = data[rr]
        return 0

    # ---- fg_deform_embed / bwd
    @staticmethod
    def _embed(p, multires):
        out = [p]
        for f in range(multires):
            out += [np.sin(p * np.float32(2.0 ** f)), np.cos(p * np.float32(2.0 ** f))]
        return np.concatenate(out, 1)

    def fg_deform_embed(self, N, x, x2, t_emb, t_ch, multires, ld, e, stream):
        self.calls.append("embed")
        if N == 0:
            return 0
        out = _arr(e, (N, ld))
        out[:] = 0
        parts = [self._embed(_arr(x, (N, 3)), multires)]
        if x2:
            parts.append(self._embed(_arr(x2, (N, 3)), multires))
        if t_ch:
            parts.append(np.broadcast_to(_arr(t_emb, (t_ch,)), (N, t_ch)))
        cat = np.concatenate(parts, 1)
        out[:, :cat.shape[1]] = cat
        return 0

    def fg_deform_embed_bwd(self, N, x, de, multires, ld, dx, stream):
        self.calls.append("embed_bwd")
        if N == 0:
            return 0
        with torch.enable_grad():  # called from inside autograd.Function.backward, where grad mode is off
            xv = torch.from_numpy(_arr(x, (N, 3)).copy()).double().requires_grad_(True)
            parts = [xv]
            for f in range(multires):
                parts += [torch.sin(xv * 2.0 ** f), torch.cos(xv * 2.0 ** f)]
            emb = torch.cat(parts, 1)
            (emb * torch.from_numpy(_arr(de, (N, ld))[:, :emb.shape[1]].copy()).double()).sum().backward()
        _arr(dx, (N, 3))[:] = xv.grad.float().numpy()
        return 0

    # ---- fg_mlp_linear
    def fg_mlp_linear(self, mode, M, n_out, a0, k0, a1, k1, w_hi, w_lo, bias, mask_in, out, mask_out, stream):
        self.calls.append(("linear", mode, M, n_out, k0, k1))
        if M == 0:
            return 0
        A = _arr(a0, (M, k0)).astype(np.float64)
        if k1:
            A = np.concatenate([A, _arr(a1, (M, k1)).astype(np.float64)], 1)
        W = _arr(w_hi, (n_out, k0 + k1)).astype(np.float64) + _arr(w_lo, (n_out, k0 + k1)).astype(np.float64)
        acc = A @ W.T
        o = _arr(out, (M, n_out))
        if mode == 0:  # FG_MLP_RELU
            acc = acc + _arr(bias, (n_out,)).astype(np.float64)
            o[:] = np.maximum(acc, 0).astype(np.float32)
            bits = (o > 0).reshape(M, n_out // 32, 32).astype(np.uint64)
            words = (bits << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
            _arr(mask_out, (M, n_out // 32), np.uint32)[:] = words
        elif mode == 1:  # FG_MLP_LINEAR
            o[:] = (acc + _arr(bias, (n_out,)).astype(np.float64)).astype(np.float32)
        else:  # FG_MLP_DGRAD
            words = _arr(mask_in, (M, n_out // 32), np.uint32).astype(np.uint64)
            keep = ((words[:, :, None] >> np.arange(32, dtype=np.uint64)) & 1).reshape(M, n_out).astype(bool)
            o[:] = np.where(keep, acc, 0).astype(np.float32)
        return 0

    # ---- fg_mlp_wgrad
    def fg_mlp_wgrad(self, N, dz, a, k_in, dw, ld_dw, col0, db, stream):
        self.calls.append(("wgrad", N, k_in))
        if N == 0:
            return 0
        DZ, A = _arr(dz, (N, 256)).astype(np.float64), _arr(a, (N, k_in)).astype(np.float64)
        _arr(dw, (256, ld_dw))[:, col0:col0 + k_in] += (DZ.T @ A).astype(np.float32)
        if db:
            _arr(db, (256,))[:] += DZ.sum(0).astype(np.float32)
        return 0

    # ---- fg_deform_apply_fwd / bwd (the formulas of freegaussian_model.py:841-845 / utils.py:137-159 through autograd)
    @staticmethod
    def _apply(head, means, scales_log, quats):
        from oracle import deform as OD

        w, v = head[:, 0:3], head[:, 3:6]
        th = w.norm(dim=-1, keepdim=True)
        T = OD.exp_se3(torch.cat([w / th + 1e-5, v / th + 1e-5], -1), th)
        mh = torch.bmm(T, torch.cat([means, torch.ones_like(means[:, :1])], -1).unsqueeze(-1)).squeeze(-1)
        return mh[:, :3] / mh[:, 3:], torch.exp(scales_log) + head[:, 10:13], quats / quats.norm(dim=-1, keepdim=True) + head[:, 6:10]

    def fg_deform_apply_fwd(self, N, head, means, scales_log, quats, mo, so, qo, stream):
        self.calls.append("apply_fwd")
        t = lambda p, w: torch.from_numpy(_arr(p, (N, w)).copy()).double()  # noqa: E731
        outs = self._apply(t(head, HEAD_LD), t(means, 3), t(scales_log, 3), t(quats, 4))
        for dst, w_, o in zip((mo, so, qo), (3, 3, 4), outs):
            _arr(dst, (N, w_))[:] = o.float().numpy()
        return 0

    def fg_deform_apply_bwd(self, N, head, means, scales_log, quats, g_m, g_s, g_q, v_head, v_m, v_s, v_q, stream):
        self.calls.append("apply_bwd")
        t = lambda p, w: torch.from_numpy(_arr(p, (N, w)).copy()).double()  # noqa: E731
        with torch.enable_grad():  # called from inside autograd.Function.backward, where grad mode is off
            ins = [t(head, HEAD_LD).requires_grad_(True), t(means, 3).requires_grad_(True), t(scales_log, 3).requires_grad_(True),
                   t(quats, 4).requires_grad_(True)]
            outs = self._apply(*ins)
            sum((o * g).sum() for o, g in zip(outs, (t(g_m, 3), t(g_s, 3), t(g_q, 4)))).backward()
        for dst, w_, i in zip((v_head, v_m, v_s, v_q), (HEAD_LD, 3, 3, 4), ins):
            _arr(dst, (N, w_))[:] = i.grad.float().numpy()
        return 0

    # ---- fg_rows_active / fg_rows_gather (csrc/rows.cu)
    def fg_rows_workspace_bytes(self, N):
        return 16

    def fg_rows_active(self, N, g, ld, idx, count_dev, ws, ws_bytes, stream):
        self.calls.append("rows_active")
        out = _arr(idx, (N,), np.int32)
        out[:] = 0
        act = np.nonzero((_arr(g, (N, ld)) != 0).any(1))[0].astype(np.int32)
        out[:act.size] = act
        ctypes.c_int64.from_address(int(count_dev)).value = int(act.size)
        return 0

    def fg_rows_gather(self, M, idx, count_dev, src, row_bytes, dst, stream):
        self.calls.append("rows_gather")
        if M == 0:
            return 0
        n = ctypes.c_int64.from_address(int(count_dev)).value
        w = row_bytes // 4
        rows = _arr(idx, (M,), np.int32)[:min(n, M)]
        big = int(rows.max()) + 1 if rows.size else 1
        s_ = _arr(src, (big, w), np.uint32)
        d = _arr(dst, (M, w), np.uint32)
        d[:] = 0
        d[:rows.size] = s_[rows]
        return 0

    # ---- time branch
    def fg_time_branch_fwd(self, t, multires, in_ch, hidden, out_ch, w1, b1, w2, b2, emb, h, out, stream):
        self.calls.append("time_fwd")
        tv = _arr(t, (1,)).astype(np.float64)
        e = [tv]
        for k in range(multires):
            e += [np.sin(tv * 2.0 ** k), np.cos(tv * 2.0 ** k)]
        e = np.concatenate(e)
        _arr(emb, (in_ch,))[:] = e.astype(np.float32)
        if not w1:
            return 0
        hv = np.maximum(_arr(w1, (hidden, in_ch)).astype(np.float64) @ e + _arr(b1, (hidden,)), 0)
        _arr(h, (hidden,))[:] = hv.astype(np.float32)
        _arr(out, (out_ch,))[:] = (_arr(w2, (out_ch, hidden)).astype(np.float64) @ hv + _arr(b2, (out_ch,))).astype(np.float32)
        return 0

    def fg_time_branch_bwd(self, in_ch, hidden, out_ch, emb, h, w2, g_out, dw1, db1, dw2, db2, stream):
        self.calls.append("time_bwd")
        e, hv, g = _arr(emb, (in_ch,)).astype(np.float64), _arr(h, (hidden,)).astype(np.float64), _arr(g_out, (out_ch,)).astype(np.float64)
        _arr(db2, (out_ch,))[:] = g.astype(np.float32)
        _arr(dw2, (out_ch, hidden))[:] = np.outer(g, hv).astype(np.float32)
        dh = (_arr(w2, (out_ch, hidden)).astype(np.float64).T @ g) * (hv > 0)
        _arr(db1, (hidden,))[:] = dh.astype(np.float32)
        _arr(dw1, (hidden, in_ch))[:] = np.outer(dh, e).astype(np.float32)
        return 0

    def fg_last_error(self):
        return b"fake"
