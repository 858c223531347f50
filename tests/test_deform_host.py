"""Host orchestration of the deformation / control networks on CPU (no GPU, no CUDA library).

`freegaussian_b200/deform.py` is driven here with `tests/fake_mlp_lib.FakeMlpLib` standing in for the C ABI: the packing plan,
the [h | embedding] operand order of the skip layer, the assembly of every parameter gradient (reference column order, head
transposes, time gradient through the bias gradients), the gradient of the embedded points and the sparse backward are all
Python, and must reproduce the oracle whatever executes the individual calls.
"""
import numpy as np
import pytest
import torch

from fake_mlp_lib import FakeMlpLib
from oracle import deform as OD
from util import grad_rel_err, rel_err


@pytest.fixture()
def D(monkeypatch):
    from freegaussian_b200 import deform

    fake = FakeMlpLib()
    monkeypatch.setattr(deform._lib, "lib", lambda: fake)
    monkeypatch.setattr(deform, "_stream", lambda: 0)
    deform._fake = fake
    deform._ACTIVE_ROWS.clear()
    return deform


def _scene(n, seed):
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(n, 3, generator=g) - 0.5) * 6.0
    scales_log = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005)
    quats = torch.randn(n, 4, generator=g)
    ws = [torch.randn(n, k, generator=g) for k in (3, 3, 4)]
    return means, scales_log, quats, ws


@pytest.mark.parametrize("is_blender", [True, False])
@pytest.mark.parametrize("sparse", [False, True])
def test_deform_orchestration_matches_oracle(D, is_blender, sparse):
    n = 150
    params = OD.init_params(is_blender=is_blender, seed=12)
    net = D.DeformNetwork(is_blender=is_blender)
    net.load_state_dict(params)
    means, scales_log, quats, ws = _scene(n, 5)
    if sparse:  # 60 % of the rows receive no gradient: the backward runs on the rest
        keep = (torch.arange(n) % 5 < 2).float()[:, None]
        ws = [w * keep for w in ws]
    t = torch.tensor([[0.35]])
    # oracle
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    lo = [x.clone().requires_grad_(True) for x in (means, scales_log, quats)]
    want = OD.deform_gaussians(P, *lo, t.expand(n, -1), is_blender)
    sum((o * w).sum() for o, w in zip(want, ws)).backward()
    # product host code over the fake library (module.head() refuses CPU tensors, so its two lines are spelled out)
    lp = [x.clone().requires_grad_(True) for x in (means, scales_log, quats)]
    spec = D._Spec(D.MLP_EMBED_LD, net.input_ch, net.multires, D._HEADS, input_grad=False)
    head = D._Trunk.apply(lp[0].detach().contiguous(), None, net._time_row(t).contiguous(), spec, *net._params())
    got = D._Apply.apply(head, *lp)
    sum((o * w).sum() for o, w in zip(got, ws)).backward()
    for a, b in zip(got, want):
        assert rel_err(a, b) < 1e-5
    for a, b in zip(lp, lo):
        assert grad_rel_err(a.grad, b.grad) < 1e-5
    for k, v in net.named_parameters():
        assert grad_rel_err(v.grad, P[k].grad) < 1e-4, k
    calls = D._fake.calls
    n_rows = [c[2] for c in calls if isinstance(c, tuple) and c[0] == "linear" and c[1] == 2]  # data-gradient calls
    assert len(n_rows) == 8 and (all(r == 60 for r in n_rows) if sparse else all(r == n for r in n_rows))


def test_control_orchestration_matches_oracle(D):
    n = 120
    params = OD.init_control_params(seed=13)
    net = D.ControlNetwork()
    net.load_state_dict(params)
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(n, 3, generator=g) - 0.5) * 6.0
    value = torch.randn(n, 3, generator=g) * 0.1
    ws = [torch.randn(n, k, generator=g) for k in (3, 4, 3)]
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo = x.clone().requires_grad_(True)
    want = OD.control_forward(P, xo, value)
    sum((o * w).sum() for o, w in zip(want, ws)).backward()
    xp = x.clone().requires_grad_(True)
    spec = D._Spec(D._CONTROL_LD, net.input_ch, net.multires, D._CONTROL_HEADS, input_grad=True)
    h = D._Trunk.apply(xp.contiguous(), value.contiguous(), None, spec, *net._params())
    got = (h[:, 0:3], h[:, 3:7], h[:, 7:10])
    sum((o * w).sum() for o, w in zip(got, ws)).backward()
    for a, b in zip(got, want):
        assert rel_err(a, b) < 1e-5
    assert grad_rel_err(xp.grad, xo.grad) < 1e-4
    for k, v in net.named_parameters():
        assert grad_rel_err(v.grad, P[k].grad) < 1e-4, k


def test_no_active_rows_gives_zero_gradients(D):
    n = 40
    net = D.DeformNetwork(is_blender=True)
    net.load_state_dict(OD.init_params(True, seed=3))
    means, scales_log, quats, _ = _scene(n, 7)
    spec = D._Spec(D.MLP_EMBED_LD, net.input_ch, net.multires, D._HEADS, input_grad=False)
    head = D._Trunk.apply(means, None, net._time_row(torch.tensor([[0.1]])).contiguous(), spec, *net._params())
    (head * 0).sum().backward()
    assert all(float(p.grad.abs().max()) == 0.0 for p in net.parameters() if p.grad is not None)
    assert np.isfinite(head.detach().numpy()).all()


def test_sparse_backward_capacity_guess_and_overflow(D, monkeypatch):
    """The second backward of a shape sizes its row buffers from the FIRST one's count (no host read in front of the
    kernels), zero-filling the rows past the real count; a count that overflows the guess is redone densely.  All exact."""
    monkeypatch.setattr(D, "_CAP_EXTRA", 0)
    n = 400
    params = OD.init_params(is_blender=True, seed=21)
    net = D.DeformNetwork(is_blender=True)
    net.load_state_dict(params)
    means, scales_log, quats, ws0 = _scene(n, 9)
    t = torch.tensor([[0.6]])
    spec = D._Spec(D.MLP_EMBED_LD, net.input_ch, net.multires, D._HEADS, input_grad=False)

    def run(active_mod):
        keep = (torch.arange(n) % 10 < active_mod).float()[:, None]
        ws = [w * keep for w in ws0]
        P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        lo = [x.clone().requires_grad_(True) for x in (means, scales_log, quats)]
        want = OD.deform_gaussians(P, *lo, t.expand(n, -1), True)
        sum((o * w).sum() for o, w in zip(want, ws)).backward()
        for p in net.parameters():
            p.grad = None
        lp = [x.clone().requires_grad_(True) for x in (means, scales_log, quats)]
        D._fake.calls.clear()
        head = D._Trunk.apply(lp[0].detach().contiguous(), None, net._time_row(t).contiguous(), spec, *net._params())
        got = D._Apply.apply(head, *lp)
        sum((o * w).sum() for o, w in zip(got, ws)).backward()
        for a, b in zip(lp, lo):
            assert grad_rel_err(a.grad, b.grad) < 1e-5
        for k, v in net.named_parameters():
            assert grad_rel_err(v.grad, P[k].grad) < 1e-4, k
        return [c[2] for c in D._fake.calls if isinstance(c, tuple) and c[0] == "linear" and c[1] == 2]

    assert run(4) == [160] * 8                    # first backward: exact count (one host read)
    assert run(3) == [256] * 8                    # 120 active rows in a 256-row capacity guessed from 160 * 1.12
    rows = run(9)                                 # 360 active rows overflow the 256-row guess: redone on every row
    assert rows[:8] == [256] * 8 and rows[8:] == [n] * 8
    assert run(9) == [n] * 8                      # 360 * 1.12 > 0.75 n: dense from the start
