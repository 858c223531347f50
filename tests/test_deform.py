"""Deformation network: oracle vs the reference's own outputs (CPU), kernels vs oracle (GPU).

Tolerances: forward and data gradient run in 3xTF32 = fp32 accuracy, so outputs are held to 1e-5 of the tensor's scale
(the north-star budget for everything that feeds the renderer is 1e-4); weight gradients (also 3xTF32, split-K with
atomic adds) are held to 5e-5 of the reference gradient's max magnitude, far inside north_star's 1e-3.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import deform as OD
from util import grad_rel_err, rel_err

GOLDEN = Path(__file__).parent / "golden"
FIXTURES = ["blender_n300", "blender_n129_hot", "real_n200"]
GRAD_STRIDE = 97  # tests/golden/make_golden_deform.py


def _sample(g):
    """Small gradients are stored whole, large ones as a strided sample (make_golden_deform.py)."""
    return g if g.numel() <= 4096 else g[::GRAD_STRIDE]


def _load(name):
    z = np.load(GOLDEN / f"deform_{name}.npz")
    isb = bool(z["is_blender"])
    params = OD.init_params(is_blender=isb, seed=int(z["seed"]), scale=float(z["scale"]))
    return z, isb, params


def _oracle_run(z, isb, params, dtype=torch.float32):
    P = {k: v.to(dtype).requires_grad_(True) for k, v in params.items()}
    m = torch.tensor(z["means"], dtype=dtype, requires_grad=True)
    s = torch.tensor(z["scales_log"], dtype=dtype, requires_grad=True)
    q = torch.tensor(z["quats"], dtype=dtype, requires_grad=True)
    t = torch.tensor([[float(z["t"])]], dtype=dtype).expand(m.shape[0], -1)
    nm, ns, nq = OD.deform_gaussians(P, m, s, q, t, isb)
    loss = ((nm * torch.tensor(z["w_means"], dtype=dtype)).sum() + (ns * torch.tensor(z["w_scales"], dtype=dtype)).sum()
            + (nq * torch.tensor(z["w_quats"], dtype=dtype)).sum())
    loss.backward()
    return P, (m, s, q), (nm, ns, nq)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_outputs(name):
    """The restatement against what the reference's own class bodies produced (fixtures made in the build container)."""
    z, isb, params = _load(name)
    P, (m, s, q), (nm, ns, nq) = _oracle_run(z, isb, params)
    t = torch.tensor([[float(z["t"])]]).expand(m.shape[0], -1)
    d_xyz, rot, scl = OD.deform_forward({k: v.detach() for k, v in P.items()}, m.detach(), t, isb)
    for got, key in ((d_xyz, "d_xyz"), (rot, "d_rotation"), (scl, "d_scaling"), (nm, "new_means"), (ns, "new_scales"), (nq, "new_quats")):
        assert rel_err(got, torch.tensor(z[key])) < 1e-6, key
    for got, key in ((m, "grad_means"), (s, "grad_scales_log"), (q, "grad_quats")):
        assert grad_rel_err(got.grad, torch.tensor(z[key])) < 1e-5, key
    for k, v in P.items():
        g = v.grad.double().flatten()
        assert grad_rel_err(_sample(g), torch.tensor(z["grad." + k]).double()) < 1e-4, k
        assert abs(float(g.norm()) - float(z["gnorm." + k])) <= 1e-4 * float(z["gnorm." + k]), k


def test_oracle_float64_agrees_with_float32():
    z, isb, params = _load("blender_n300")
    _, _, out32 = _oracle_run(z, isb, params, torch.float32)
    _, _, out64 = _oracle_run(z, isb, params, torch.float64)
    for a, b in zip(out32, out64):
        assert rel_err(a, b) < 2e-6


def test_control_oracle_matches_reference_outputs():
    z = np.load(GOLDEN / "control_n250.npz")
    P = {k: v.requires_grad_(True) for k, v in OD.init_control_params(seed=int(z["seed"])).items()}
    x = torch.tensor(z["x"], requires_grad=True)
    d_xyz, d_rot, d_scale = OD.control_forward(P, x, torch.tensor(z["value"]))
    for got, key in ((d_xyz, "d_xyz"), (d_rot, "d_rot"), (d_scale, "d_scale")):
        assert rel_err(got, torch.tensor(z[key])) < 1e-6, key
    ((d_xyz * torch.tensor(z["w_xyz"])).sum() + (d_rot * torch.tensor(z["w_rot"])).sum() + (d_scale * torch.tensor(z["w_scale"])).sum()).backward()
    assert grad_rel_err(x.grad, torch.tensor(z["grad_x"])) < 1e-5
    for k, v in P.items():
        assert grad_rel_err(_sample(v.grad.double().flatten()), torch.tensor(z["grad." + k]).double()) < 1e-4, k


def test_module_mirrors_reference_state_dict():
    """Parameter names and shapes are the reference's (a reference checkpoint loads unchanged)."""
    from freegaussian_b200.deform import DeformNetwork

    for isb in (True, False):
        net = DeformNetwork(is_blender=isb)
        ref = OD.init_params(is_blender=isb)
        assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.items()}
        net.load_state_dict(ref, strict=True)
    with pytest.raises(RuntimeError):
        DeformNetwork(is_blender=True).head(torch.zeros(4, 3), torch.zeros(4, 1))  # no CPU path
    with pytest.raises(ValueError, match="ONE time value"):
        DeformNetwork(is_blender=True)._time_row(torch.rand(4, 1))  # per-row times are not the reference's call
    with pytest.raises(RuntimeError):  # the time branch is a kernel too (fg_time_branch_fwd): no CPU path
        DeformNetwork(is_blender=True)._time_row(torch.tensor([[0.5]]).expand(4, -1))
    from freegaussian_b200.deform import ControlNetwork

    net = ControlNetwork()
    ref = OD.init_control_params()
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.items()}
    net.load_state_dict(ref, strict=True)
    with pytest.raises(RuntimeError):
        net(torch.zeros(4, 3), torch.zeros(4, 3))


@pytest.mark.parametrize("kind", ["deform", "control"])
def test_pack_plan_rebuilds_the_layers(kind):
    """Host logic of the weight packing (column ranges, [h | embedding] reorder, transposes), applied with numpy: the packed
    operands must reproduce every layer of the reference network and the transposes its data gradients."""
    from freegaussian_b200 import deform as D

    if kind == "deform":
        params = OD.init_params(is_blender=True, seed=2)
        spec = D._Spec(96, 93, 10, D._HEADS, input_grad=False)
        heads = D._HEADS
    else:
        params = OD.init_control_params(seed=2)
        spec = D._Spec(128, 126, 10, D._CONTROL_HEADS, input_grad=True)
        heads = D._CONTROL_HEADS
    flat = []
    for i in range(8):
        flat += [params[f"linear.{i}.weight"].numpy(), params[f"linear.{i}.bias"].numpy()]
    for name, _ in heads:
        flat += [params[name + ".weight"].numpy(), params[name + ".bias"].numpy()]
    bufs = {k: np.zeros(shape, np.float32) for k, shape in D._packed_shapes(spec).items()}
    plan = D._pack_plan(spec)
    assert len(plan) <= 32
    for p, col0, cols, dst, dst_col0, transpose, row0 in plan:
        block = flat[p][:, col0:col0 + cols]
        if transpose:
            block = block.T
        bufs[dst][row0:row0 + block.shape[0], dst_col0:dst_col0 + block.shape[1]] = block
    rng = np.random.default_rng(0)
    emb = np.zeros((5, spec.emb_ld), np.float32)
    emb[:, :spec.emb_ch] = rng.standard_normal((5, spec.emb_ch))
    h = rng.standard_normal((5, 256)).astype(np.float32)
    for i in range(8):
        W = flat[2 * i]
        if i == 0:
            want, got = emb[:, :spec.emb_ch] @ W.T, emb @ bufs[("w", 0)].T
        elif i == 5:
            want = np.concatenate([emb[:, :spec.emb_ch], h], 1) @ W.T          # reference order [embedding | h]
            got = np.concatenate([h, emb], 1) @ bufs[("w", 5)].T               # operand order [h | embedding]
        else:
            want, got = h @ W.T, h @ bufs[("w", i)].T
        assert np.allclose(got, want, atol=1e-5), i
        if i > 0:  # data gradient with respect to the hidden input: dz . W[:, hidden columns]
            dz = rng.standard_normal((5, 256)).astype(np.float32)
            Wh = W[:, spec.emb_ch:] if i == 5 else W
            assert np.allclose(dz @ bufs[("wt", i)].T, dz @ Wh, atol=1e-5), i
    Whead = np.concatenate([flat[16 + 2 * j] for j in range(len(heads))], 0)
    n_out = Whead.shape[0]
    assert np.allclose((h @ bufs[("w_head",)].T)[:, :n_out], h @ Whead.T, atol=1e-5)
    assert np.all(bufs[("w_head",)][n_out:] == 0)
    g = rng.standard_normal((5, 32)).astype(np.float32)
    assert np.allclose(g @ bufs[("wt_head",)].T, g[:, :n_out] @ Whead, atol=1e-5)
    if spec.input_grad:  # d/d(embedding) = [dz_0 | dz_5] . wt_emb^T
        dz0, dz5 = rng.standard_normal((5, 256)).astype(np.float32), rng.standard_normal((5, 256)).astype(np.float32)
        want = dz0 @ flat[0] + dz5 @ flat[10][:, :spec.emb_ch]
        got = np.concatenate([dz0, dz5], 1) @ bufs[("wt_emb",)].T
        assert np.allclose(got[:, :spec.emb_ch], want, atol=1e-4) and np.all(got[:, spec.emb_ch:] == 0)


# ------------------------------------------------------------------------------------------------------- GPU
def _hilo(x):
    """Host model of the operand split (round to nearest tf32, ties away): hi, lo."""
    b = x.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    b = (b + 0x1000) & 0xFFFFE000
    b = torch.where(b >= 2 ** 31, b - 2 ** 32, b).to(torch.int32)
    hi = b.view(torch.float32)
    return hi, x - hi


def _unpack_bits(bits, n_cols):
    """[M, n_cols/32] int32 words -> bool [M, n_cols]."""
    b = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    return ((b.unsqueeze(-1) >> torch.arange(32)) & 1).reshape(bits.shape[0], n_cols).bool()


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1, 127, 128, 129, 1000, 148 * 128 * 2 + 77])
def test_linear_relu_is_fp32_accurate(built_lib, M):
    from freegaussian_b200 import _lib
    from freegaussian_b200.deform import _linear

    g = torch.Generator().manual_seed(M)
    for k0, k1 in ((96, 0), (256, 0), (256, 96)):
        a = torch.randn(M, k0 + k1, generator=g)
        w = torch.randn(256, k0 + k1, generator=g) / (k0 + k1) ** 0.5
        bias = torch.randn(256, generator=g)
        pre = a.double() @ w.double().T + bias.double()
        ad, wd, bd = a.cuda(), w.cuda(), bias.cuda()
        a0, a1 = ad[:, :k0].contiguous(), ad[:, k0:].contiguous()
        out = torch.empty(M, 256, device="cuda")
        bits = torch.empty(M, 8, dtype=torch.int32, device="cuda")
        _linear(_lib.MLP_RELU, M, 256, a0, k0, a1 if k1 else None, k1, _hilo(wd), bd, None, out, bits)
        assert rel_err(out, torch.relu(pre)) < 3e-6, (M, k0, k1)
        # the mask the data-gradient pass reads is exactly (output > 0)
        assert bool((_unpack_bits(bits, 256) == (out.cpu() > 0)).all())


@pytest.mark.gpu
@pytest.mark.parametrize("M", [5, 4096 + 3])
def test_linear_head_and_dgrad_modes(built_lib, M):
    from freegaussian_b200 import _lib
    from freegaussian_b200.deform import _linear

    g = torch.Generator().manual_seed(M)
    a = torch.randn(M, 256, generator=g)
    w = torch.randn(32, 256, generator=g) / 16
    bias = torch.randn(32, generator=g)
    out = torch.empty(M, 32, device="cuda")
    ad, wd, bd = a.cuda(), w.cuda(), bias.cuda()
    _linear(_lib.MLP_LINEAR, M, 32, ad, 256, None, 0, _hilo(wd), bd, None, out, None)
    assert rel_err(out, a.double() @ w.double().T + bias.double()) < 3e-6
    # data gradient: dz_prev = (dz . Wt^T) where the ReLU input was positive
    for k in (32, 256):
        dz = torch.randn(M, k, generator=g)
        wt = torch.randn(256, k, generator=g) / k ** 0.5
        keep = torch.rand(M, 256, generator=g) > 0.5
        words = (keep.reshape(M, 8, 32).to(torch.int64) << torch.arange(32)).sum(-1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).cuda()
        got = torch.empty(M, 256, device="cuda")
        dzd, wtd = dz.cuda(), wt.cuda()
        _linear(_lib.MLP_DGRAD, M, 256, dzd, k, None, 0, _hilo(wtd), None, words, got, None)
        want = (dz.double() @ wt.double().T) * keep
        assert grad_rel_err(got, want) < 3e-6
        assert bool(((got.cpu() == 0) | keep).all())


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 15, 16, 1000, 100_003])
def test_weight_gradient_kernel(built_lib, N):
    """dW += dz^T a and db += column sums, 3xTF32: fp32-accurate, additive, every row counted once."""
    from freegaussian_b200 import _lib
    from freegaussian_b200._lib import check, ptr

    g = torch.Generator().manual_seed(N)
    st = torch.cuda.current_stream().cuda_stream
    for k_in in (256, 128, 96, 32):
        dz = torch.randn(N, 256, generator=g)
        a = torch.randn(N, k_in, generator=g)
        dzd, ad = dz.cuda(), a.cuda()
        dw = torch.zeros(256, k_in, device="cuda")
        db = torch.zeros(256, device="cuda")
        check(_lib.lib().fg_mlp_wgrad(N, ptr(dzd), ptr(ad), k_in, ptr(dw), k_in, 0, ptr(db), st))
        want_w, want_b = dz.double().T @ a.double(), dz.double().sum(0)
        # fp32 accumulation over up to 1e5 rows (the reference's fp32 GEMM rounds the same way): 2e-5 of the largest entry
        assert grad_rel_err(dw, want_w) < 2e-5, (N, k_in)
        assert grad_rel_err(db, want_b) < 2e-5, (N, k_in)
        check(_lib.lib().fg_mlp_wgrad(N, ptr(dzd), ptr(ad), k_in, ptr(dw), k_in, 0, None, st))  # adds; db untouched
        assert grad_rel_err(dw, 2 * want_w) < 2e-5
        assert grad_rel_err(db, want_b) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("t_ch,multires,two,ld", [(30, 10, False, 96), (21, 10, False, 96), (0, 4, False, 96), (0, 10, True, 128)])
def test_embedding_matches_reference_layout(built_lib, t_ch, multires, two, ld):
    from freegaussian_b200 import _lib
    from freegaussian_b200._lib import check, ptr

    g = torch.Generator().manual_seed(1)
    n = 1001
    x = (torch.rand(n, 3, generator=g) - 0.5) * 6.0
    x2 = torch.randn(n, 3, generator=g) * 0.1
    t_emb = torch.randn(t_ch, generator=g)
    e = torch.empty(n, ld, device="cuda")
    xd, x2d, td = x.cuda(), x2.cuda(), t_emb.cuda()
    st = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().fg_deform_embed(n, ptr(xd), ptr(x2d) if two else None, ptr(td) if t_ch else None, t_ch, multires, ld, ptr(e), st))
    parts = [OD.embed(x, multires)] + ([OD.embed(x2, multires)] if two else []) + [t_emb.expand(n, -1)]
    want = torch.cat(parts, -1)
    got = e.cpu()
    worst = float((got[:, :want.shape[1]] - want).abs().max())
    assert worst < 1e-6, worst  # sin / cos of arguments up to 3 * 2^9, same fp32 argument on both sides
    assert float(got[:, want.shape[1]:].abs().max()) == 0.0
    # VJP with respect to x against autograd
    de = torch.randn(n, ld, generator=g)
    xr = x.clone().double().requires_grad_(True)
    (OD.embed(xr, multires) * de[:, :3 + 6 * multires].double()).sum().backward()
    dx = torch.empty(n, 3, device="cuda")
    ded = de.cuda()
    check(_lib.lib().fg_deform_embed_bwd(n, ptr(xd), ptr(ded), multires, ld, ptr(dx), st))
    assert grad_rel_err(dx, xr.grad) < 1e-5


@pytest.mark.gpu
def test_embedding_gradient_gemm(built_lib):
    """FG_MLP_LINEAR with 128 outputs over two activation sources (K = 512): the gradient of the embedding."""
    from freegaussian_b200 import _lib
    from freegaussian_b200.deform import _linear

    g = torch.Generator().manual_seed(4)
    M = 3001
    a0, a1 = torch.randn(M, 256, generator=g), torch.randn(M, 256, generator=g)
    w = torch.randn(128, 512, generator=g) / 22
    out = torch.empty(M, 128, device="cuda")
    a0d, a1d, wd = a0.cuda(), a1.cuda(), w.cuda()
    _linear(_lib.MLP_LINEAR, M, 128, a0d, 256, a1d, 256, _hilo(wd), torch.zeros(128, device="cuda"), None, out, None)
    assert rel_err(out, torch.cat([a0, a1], 1).double() @ w.double().T) < 6e-6  # fp32 accumulation over K = 512


@pytest.mark.gpu
def test_apply_kernels_match_autograd(built_lib):
    from freegaussian_b200.deform import _Apply

    g = torch.Generator().manual_seed(2)
    n = 777
    head = torch.zeros(n, 32)
    head[:, :13] = torch.randn(n, 13, generator=g) * 0.3
    m, s, q = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g) - 3, torch.randn(n, 4, generator=g)
    ws = [torch.randn(n, k, generator=g) for k in (3, 3, 4)]

    def ref(head, m, s, q):
        w, v = head[:, 0:3], head[:, 3:6]
        th = w.norm(dim=-1, keepdim=True)
        T = OD.exp_se3(torch.cat([w / th + 1e-5, v / th + 1e-5], -1), th)
        mh = torch.bmm(T, torch.cat([m, torch.ones_like(m[:, :1])], -1).unsqueeze(-1)).squeeze(-1)
        return mh[:, :3] / mh[:, 3:], torch.exp(s) + head[:, 10:13], q / q.norm(dim=-1, keepdim=True) + head[:, 6:10]

    cpu = [t.clone().double().requires_grad_(True) for t in (head, m, s, q)]
    outs = ref(*cpu)
    sum((o * w.double()).sum() for o, w in zip(outs, ws)).backward()
    dev = [t.clone().cuda().requires_grad_(True) for t in (head, m, s, q)]
    got = _Apply.apply(*dev)
    sum((o * w.cuda()).sum() for o, w in zip(got, ws)).backward()
    for a, b in zip(got, outs):
        assert rel_err(a, b) < 2e-6
    for a, b in zip(dev, cpu):
        assert grad_rel_err(a.grad[:, :13] if a.shape[1] == 32 else a.grad, b.grad[:, :13] if b.shape[1] == 32 else b.grad) < 1e-5
    assert float(dev[0].grad[:, 13:].abs().max()) == 0.0


def _net_from(params, isb):
    from freegaussian_b200.deform import DeformNetwork

    net = DeformNetwork(is_blender=isb)
    net.load_state_dict(params, strict=True)
    return net.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_network_matches_reference_fixture(built_lib, name):
    """End to end against the outputs and gradients of the reference's own code (tests/golden/deform_*.npz)."""
    z, isb, params = _load(name)
    net = _net_from(params, isb)
    m, s, q = (torch.tensor(z[k]).cuda().requires_grad_(True) for k in ("means", "scales_log", "quats"))
    t = torch.tensor([[float(z["t"])]]).cuda().expand(m.shape[0], -1)
    nm, ns, nq = net.deform_gaussians(m, s, q, t)
    for got, key in ((nm, "new_means"), (ns, "new_scales"), (nq, "new_quats")):
        assert rel_err(got, torch.tensor(z[key])) < 1e-5, key
    loss = (nm * torch.tensor(z["w_means"]).cuda()).sum() + (ns * torch.tensor(z["w_scales"]).cuda()).sum() + \
        (nq * torch.tensor(z["w_quats"]).cuda()).sum()
    loss.backward()
    for got, key in ((m, "grad_means"), (s, "grad_scales_log"), (q, "grad_quats")):
        assert grad_rel_err(got.grad, torch.tensor(z[key])) < 1e-5, key
    # The "hot" fixture (weights x2: activations grow ~2^8 through the trunk) has ReLU inputs that sit within fp32
    # rounding of zero; any fp32 implementation with another summation order flips a few of them, and with 129 rows one
    # flipped unit moves a gradient entry by ~1e-4.  It is held to north_star's 1e-3, the others to 5e-5.
    tol = 1e-3 if "hot" in name else 5e-5
    for k, v in net.named_parameters():
        assert grad_rel_err(_sample(v.grad.double().flatten().cpu()), torch.tensor(z["grad." + k]).double()) < tol, k
        assert abs(float(v.grad.double().norm()) - float(z["gnorm." + k])) <= tol * float(z["gnorm." + k]), k
    # the reference-signature forward: same d_xyz / rotation / scaling
    d_xyz, rot, scl = net(m.detach(), t)
    for got, key in ((d_xyz, "d_xyz"), (rot, "d_rotation"), (scl, "d_scaling")):
        assert rel_err(got, torch.tensor(z[key])) < 1e-5, key


@pytest.mark.gpu
def test_control_network_matches_reference_fixture(built_lib):
    """Stage-2 network against the outputs and gradients of the reference's own class, including d(loss)/dx."""
    from freegaussian_b200.deform import ControlNetwork

    z = np.load(GOLDEN / "control_n250.npz")
    net = ControlNetwork()
    net.load_state_dict(OD.init_control_params(seed=int(z["seed"])), strict=True)
    net = net.cuda()
    x = torch.tensor(z["x"]).cuda().requires_grad_(True)
    outs = net(x, torch.tensor(z["value"]).cuda())
    for got, key in zip(outs, ("d_xyz", "d_rot", "d_scale")):
        assert rel_err(got, torch.tensor(z[key])) < 1e-5, key
    sum((o * torch.tensor(z[k]).cuda()).sum() for o, k in zip(outs, ("w_xyz", "w_rot", "w_scale"))).backward()
    assert grad_rel_err(x.grad, torch.tensor(z["grad_x"])) < 5e-5
    for k, v in net.named_parameters():
        assert grad_rel_err(_sample(v.grad.double().flatten().cpu()), torch.tensor(z["grad." + k]).double()) < 5e-5, k
    # rows without a gradient are skipped and get an exactly-zero d(loss)/dx
    x2 = torch.tensor(z["x"]).cuda().requires_grad_(True)
    keep = (torch.arange(x2.shape[0], device="cuda") % 3 == 0).float()[:, None]
    sum((o * torch.tensor(z[k]).cuda() * keep).sum() for o, k in zip(net(x2, torch.tensor(z["value"]).cuda()), ("w_xyz", "w_rot", "w_scale"))).backward()
    assert float(x2.grad[keep[:, 0] == 0].abs().max()) == 0.0
    assert grad_rel_err(x2.grad[keep[:, 0] == 1], torch.tensor(z["grad_x"]).cuda()[keep[:, 0] == 1]) < 5e-5


@pytest.mark.gpu
def test_sparse_backward_equals_dense(built_lib):
    """Rows with a zero incoming gradient (culled Gaussians) are skipped by the backward: same gradients as the dense pass."""
    from freegaussian_b200 import deform as D

    params = OD.init_params(is_blender=True, seed=21)
    net = _net_from(params, True)
    g = torch.Generator().manual_seed(9)
    n = 5000
    m = ((torch.rand(n, 3, generator=g) - 0.5) * 6.0).cuda()
    s = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005).cuda()
    q = torch.randn(n, 4, generator=g).cuda()
    t = torch.tensor([[0.6]]).cuda().expand(n, -1)
    seen = (torch.rand(n, 1, generator=g) < 0.3).float().cuda()  # 30 % of the Gaussians receive a gradient
    ws = [torch.randn(n, k, generator=g).cuda() * seen for k in (3, 3, 4)]

    def run(sparse):
        D.SPARSE_BACKWARD = sparse
        try:
            net.zero_grad()
            leaves = [x.clone().requires_grad_(True) for x in (m, s, q)]
            outs = net.deform_gaussians(*leaves, t)
            sum((o * w).sum() for o, w in zip(outs, ws)).backward()
            return [p.grad.clone() for p in net.parameters()] + [x.grad for x in leaves]
        finally:
            D.SPARSE_BACKWARD = True

    for a, b in zip(run(True), run(False)):
        assert grad_rel_err(a, b) < 1e-5
    # nothing visible at all: every gradient is exactly zero
    ws = [w * 0 for w in ws]
    assert all(float(x.abs().max()) == 0.0 for x in run(True)[:-3])


@pytest.mark.gpu
def test_network_then_render_matches_chained_oracles(built_lib):
    """The stage-1 forward/backward as the model runs it (freegaussian_model.py:836-868): network -> rasterization -> loss,
    kernels against the two oracles chained with torch.autograd.  Image tolerance 1e-4, gradient tolerance 1e-3 (north_star)."""
    from freegaussian_b200.rendering import rasterization
    from oracle import render as OR
    from util import small_scene

    sc = small_scene(700, 80, 48, views=1, seed=3)
    W, H = 80, 48
    params = OD.init_params(is_blender=True, seed=31, scale=0.5)  # small motions: the scene stays in view
    net = _net_from(params, True)
    g = torch.Generator().manual_seed(8)
    w_img = torch.rand(1, H, W, 4, generator=g)
    t = torch.tensor([[0.45]])
    leaves_cpu = [sc.means.clone().requires_grad_(True), torch.log(sc.scales).requires_grad_(True), sc.quats.clone().requires_grad_(True)]
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    m2, s2, q2 = OD.deform_gaussians(P, *leaves_cpu, t.expand(700, -1), True)
    ro, ao, _ = OR.rasterization(m2, q2, s2, sc.opacities, sc.sh, sc.viewmats, sc.Ks, W, H, render_mode="RGB+ED", sh_degree=3)
    (ro * w_img).sum().backward()
    d = sc.to("cuda")
    leaves = [d.means.clone().requires_grad_(True), torch.log(d.scales).requires_grad_(True), d.quats.clone().requires_grad_(True)]
    gm2, gs2, gq2 = net.deform_gaussians(*leaves, t.cuda().expand(700, -1))
    r, a, _ = rasterization(gm2, gq2, gs2, d.opacities, d.sh, d.viewmats, d.Ks, W, H, render_mode="RGB+ED", sh_degree=3)
    (r * w_img.cuda()).sum().backward()
    assert rel_err(r, ro) < 1e-4 and rel_err(a, ao) < 1e-4
    for got, want in zip(leaves, leaves_cpu):
        assert grad_rel_err(got.grad, want.grad) < 1e-3
    for k, v in net.named_parameters():
        assert grad_rel_err(v.grad, P[k].grad) < 1e-3, k


@pytest.mark.gpu
def test_full_size_rows_are_independent(built_lib):
    """cfg3 size (1 M Gaussians): every row depends on its own Gaussian only, so a random sample of rows must equal
    the oracle evaluated on just those rows; and the weight gradient is linear in the row set (two halves sum to the whole)."""
    isb = True
    params = OD.init_params(is_blender=isb, seed=11)
    net = _net_from(params, isb)
    n = 1_000_000
    g = torch.Generator().manual_seed(5)
    m = (torch.rand(n, 3, generator=g) - 0.5) * 6.0
    s = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005)
    q = torch.randn(n, 4, generator=g)
    t = torch.tensor([[0.25]])
    md, sd, qd = m.cuda(), s.cuda(), q.cuda()
    with torch.no_grad():
        nm, ns, nq = net.deform_gaussians(md, sd, qd, t.cuda().expand(n, -1))
    idx = torch.randperm(n, generator=g)[:3000]
    om, os_, oq = OD.deform_gaussians(params, m[idx], s[idx], q[idx], t.expand(len(idx), -1), isb)
    assert rel_err(nm[idx.cuda()], om) < 1e-5 and rel_err(ns[idx.cuda()], os_) < 1e-5 and rel_err(nq[idx.cuda()], oq) < 1e-5
    assert bool(torch.isfinite(nm).all())

    def wgrad(lo, hi):
        net.zero_grad()
        a, b, c = net.deform_gaussians(md[lo:hi], sd[lo:hi], qd[lo:hi], t.cuda().expand(hi - lo, -1))
        (a.sum() + b.sum() + c.sum()).backward()
        return torch.cat([p.grad.flatten() for p in net.parameters()]).double()

    k = 200_000
    whole, first, second = wgrad(0, k), wgrad(0, k // 2 + 13), wgrad(k // 2 + 13, k)
    assert grad_rel_err(first + second, whole) < 1e-3
