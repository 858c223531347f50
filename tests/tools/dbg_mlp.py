#!/usr/bin/env python
"""Stage-by-stage check of the deformation-network kernels on a GPU (prints errors, asserts nothing).

    python tests/tools/dbg_mlp.py linear|modes|wgrad|aux|e2e|layer|perf

Lives under tests/ because it checks against the oracle and the golden fixtures (test infrastructure).

Each stage is a separate process on purpose: a trapped kernel poisons the CUDA context.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from freegaussian_b200 import _lib  # noqa: E402
from freegaussian_b200.deform import DeformNetwork, _linear  # noqa: E402
from oracle import deform as OD  # noqa: E402
from test_deform import _hilo  # noqa: E402


def err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(1e-30, float(b.abs().max())))


def stage_linear():
    g = torch.Generator().manual_seed(0)
    for M, k0, k1 in ((128, 96, 0), (128, 256, 0), (300, 256, 96), (40000, 256, 0)):
        a = torch.randn(M, k0 + k1, generator=g)
        w = torch.randn(256, k0 + k1, generator=g) / (k0 + k1) ** 0.5
        bias = torch.randn(256, generator=g)
        ad, wd, bd = a.cuda(), w.cuda(), bias.cuda()
        a0, a1 = ad[:, :k0].contiguous(), ad[:, k0:].contiguous()
        out = torch.zeros(M, 256, device="cuda")
        bits = torch.zeros(M, 8, dtype=torch.int32, device="cuda")
        _linear(_lib.MLP_RELU, M, 256, a0, k0, a1 if k1 else None, k1, _hilo(wd), bd, None, out, bits)
        torch.cuda.synchronize()
        want = torch.relu(a.double() @ w.double().T + bias.double())
        print(f"relu M={M} k=({k0},{k1}): rel err {err(out, want):.3e}", flush=True)
        if err(out, want) > 1e-3:
            got = out.double().cpu()
            print("   got[0,:8]", got[0, :8].tolist(), "\n   want[0,:8]", want[0, :8].tolist())
            print("   rows err", [(r, err(got[r], want[r])) for r in (0, 1, 7, 8, 31, 32, 64, 127) if r < M])
            print("   cols err", [(c, err(got[:, c], want[:, c])) for c in (0, 1, 7, 8, 31, 32, 128, 255)])


def stage_wgrad():
    from freegaussian_b200._lib import check, ptr

    g = torch.Generator().manual_seed(3)
    st = torch.cuda.current_stream().cuda_stream
    for N, k_in in ((16, 256), (1000, 256), (1000, 96), (100_003, 256)):
        dz = torch.randn(N, 256, generator=g)
        a = torch.randn(N, k_in, generator=g)
        dzd, ad = dz.cuda(), a.cuda()
        dw, db = torch.zeros(256, k_in, device="cuda"), torch.zeros(256, device="cuda")
        check(_lib.lib().fg_mlp_wgrad(N, ptr(dzd), ptr(ad), k_in, ptr(dw), k_in, 0, ptr(db), st))
        torch.cuda.synchronize()
        want = dz.double().T @ a.double()
        print(f"wgrad N={N} k_in={k_in}: dW rel err {err(dw, want):.3e}  db rel err {err(db, dz.double().sum(0)):.3e}", flush=True)
        if err(dw, want) > 1e-3:
            got = dw.double().cpu()
            print("   got[0,:6]", got[0, :6].tolist(), "\n   want[0,:6]", want[0, :6].tolist())
            print("   got^T match?", err(got, want.T) if k_in == 256 else None)
            print("   row blocks err", [(r, f"{err(got[r:r + 32], want[r:r + 32]):.1e}") for r in range(0, 256, 32)])
            print("   col blocks err", [(c, f"{err(got[:, c:c + 32], want[:, c:c + 32]):.1e}") for c in range(0, k_in, 32)])
            for nn in (8, 16):
                print(f"   vs first {nn} rows only:", err(got, dz[:nn].double().T @ a[:nn].double()))


def stage_modes():
    import pytest

    sys.exit(pytest.main(["-q", os.path.join(ROOT, "tests/test_deform.py"), "-m", "gpu", "-k", "head_and_dgrad"]))


def stage_aux():
    import pytest

    sys.exit(pytest.main(["-q", "-x", os.path.join(ROOT, "tests/test_deform.py"), "-m", "gpu", "-k", "embedding or apply"]))


def stage_e2e():
    for name in ("blender_n300", "blender_n129_hot", "real_n200"):
        z = np.load(os.path.join(ROOT, "tests/golden", f"deform_{name}.npz"))
        isb = bool(z["is_blender"])
        params = OD.init_params(is_blender=isb, seed=int(z["seed"]), scale=float(z["scale"]))
        net = DeformNetwork(is_blender=isb)
        net.load_state_dict(params)
        net = net.cuda()
        m, s, q = (torch.tensor(z[k]).cuda().requires_grad_(True) for k in ("means", "scales_log", "quats"))
        t = torch.tensor([[float(z["t"])]]).cuda().expand(m.shape[0], -1)
        nm, ns, nq = net.deform_gaussians(m, s, q, t)
        print(name, "fwd", [f"{err(a, torch.tensor(z[k])):.2e}" for a, k in ((nm, "new_means"), (ns, "new_scales"), (nq, "new_quats"))], flush=True)
        loss = (nm * torch.tensor(z["w_means"]).cuda()).sum() + (ns * torch.tensor(z["w_scales"]).cuda()).sum() + \
            (nq * torch.tensor(z["w_quats"]).cuda()).sum()
        loss.backward()
        print("  in-grads", [f"{err(a.grad, torch.tensor(z[k])):.2e}" for a, k in ((m, "grad_means"), (s, "grad_scales_log"), (q, "grad_quats"))])
        worst = {}
        for k, v in net.named_parameters():
            gg = v.grad.double().flatten()
            worst[k] = err(gg if gg.numel() <= 4096 else gg[::97], torch.tensor(z["grad." + k]).double())
        print("  weight grads worst", max(worst.values()), max(worst, key=worst.get), flush=True)
        print("  per param", {k: f"{v:.1e}" for k, v in worst.items() if "linear" in k}, flush=True)


def stage_layer():
    """Time of one 256 -> 256 layer at 1 M rows (forward / data gradient / weight gradient)."""
    from freegaussian_b200._lib import check, ptr

    n = 1_000_000
    g = torch.Generator().manual_seed(0)
    a = torch.randn(n, 256, generator=g).cuda()
    w = _hilo((torch.randn(256, 256, generator=g) / 16).cuda())
    bias = torch.zeros(256, device="cuda")
    out = torch.empty(n, 256, device="cuda")
    bits = torch.empty(n, 8, dtype=torch.int32, device="cuda")
    dw, db = torch.zeros(256, 256, device="cuda"), torch.zeros(256, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    runs = {
        "relu": lambda: _linear(_lib.MLP_RELU, n, 256, a, 256, None, 0, w, bias, None, out, bits),
        "dgrad": lambda: _linear(_lib.MLP_DGRAD, n, 256, a, 256, None, 0, w, None, bits, out, None),
        "wgrad": lambda: check(_lib.lib().fg_mlp_wgrad(n, ptr(a), ptr(out), 256, ptr(dw), 256, 0, ptr(db), st)),
    }
    for name, fn in runs.items():
        ts = []
        for it in range(6):
            e0, e1 = ev(), ev()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        us = min(ts[2:]) * 1e3
        print(f"{name:6s}: {us:7.1f} us  ({3 * 2 * n * 256 * 256 / us / 1e6:.0f} TFLOP/s of tf32 MMA issued, {2 * n * 1024 / us / 1e3:.0f} GB/s)", flush=True)


def stage_perf():
    n = int(os.environ.get("N", 1_000_000))
    net = DeformNetwork(is_blender=True)
    net.load_state_dict(OD.init_params(True, seed=1))
    net = net.cuda()
    g = torch.Generator().manual_seed(0)
    m = ((torch.rand(n, 3, generator=g) - 0.5) * 6).cuda().requires_grad_(True)
    s = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005).cuda().requires_grad_(True)
    q = torch.randn(n, 4, generator=g).cuda().requires_grad_(True)
    t = torch.tensor([[0.3]]).cuda().expand(n, -1)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for it in range(4):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        nm, ns, nq = net.deform_gaussians(m, s, q, t)
        e1.record()
        (nm.sum() + ns.sum() + nq.sum()).backward()
        e2.record()
        torch.cuda.synchronize()
        print(f"N={n} fwd {e0.elapsed_time(e1):.3f} ms  bwd {e1.elapsed_time(e2):.3f} ms", flush=True)
    flop_fwd = 2.0 * n * (96 * 256 + 6 * 256 * 256 + 352 * 256 + 256 * 32) * 3
    print(f"fwd tensor-core work (3 products): {flop_fwd / 1e12:.2f} TFLOP -> {flop_fwd / (e0.elapsed_time(e1) * 1e-3) / 1e12:.0f} TFLOP/s")
    # the same network with torch ops (what the reference runs: fp32 nn.Linear on the GPU)
    P = {k: v.cuda() for k, v in OD.init_params(True, seed=1).items()}

    def torch_path():
        x = m.detach()
        t_emb = net._time_row(t).expand(n, -1)
        x_emb = torch.cat([x] + [f(x * (2.0 ** k)) for k in range(10) for f in (torch.sin, torch.cos)], -1)
        h = torch.cat([x_emb, t_emb], -1)
        for i in range(8):
            h = torch.relu(torch.nn.functional.linear(h, P[f"linear.{i}.weight"], P[f"linear.{i}.bias"]))
            if i == 4:
                h = torch.cat([x_emb, t_emb, h], -1)
        return h

    with torch.no_grad():
        for it in range(3):
            e0, e1 = ev(), ev()
            e0.record()
            torch_path()
            e1.record()
            torch.cuda.synchronize()
            print(f"torch fp32 trunk forward (no grad): {e0.elapsed_time(e1):.3f} ms", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    {"linear": stage_linear, "modes": stage_modes, "aux": stage_aux, "e2e": stage_e2e, "perf": stage_perf, "layer": stage_layer, "wgrad": stage_wgrad}[sys.argv[1]]()
    print(f"[{sys.argv[1]} done in {time.time() - t0:.1f}s]")
