"""Fused Adam over the Gaussian parameter groups: the restated update is pinned to
torch.optim.Adam on CPU; the kernel is compared with torch.optim.Adam on GPU."""
import pytest
import torch

from oracle import optim as OO

LRS = {"means": 1.6e-4 * 5, "features_dc": 0.0025, "features_rest": 0.0025 / 20, "opacities": 0.05,
       "scales": 0.001 * 5, "quats": 0.001}  # freegaussian_config.py:48-75


def _groups(n, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = {"means": (n, 3), "features_dc": (n, 3), "features_rest": (n, 15, 3), "opacities": (n, 1),
              "scales": (n, 3), "quats": (n, 4)}
    params = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
    grads = [{k: torch.randn(s, generator=g) * (10.0 ** torch.randint(-6, 1, (1,), generator=g).item())
              for k, s in shapes.items()} for _ in range(12)]
    for gr in grads:  # Gaussians that were not visible have exactly zero gradients
        for k in gr:
            gr[k][::3] = 0
    return params, grads


def test_restated_update_is_torch_adam():
    params, grads = _groups(257, 0)
    ref, ref_state = OO.adam_reference(params, grads, LRS)
    for k in params:
        p, m, v = params[k].clone(), torch.zeros_like(params[k]), torch.zeros_like(params[k])
        for t, g in enumerate(grads, 1):
            OO.adam_step_restated(p, g[k], m, v, t, LRS[k])
        assert torch.equal(p, ref[k]) and torch.equal(m, ref_state[k][0]) and torch.equal(v, ref_state[k][1]), k


def test_lr_schedule_endpoints_and_shards():
    from freegaussian_b200.optim import MEANS_LR_FINAL, MEANS_LR_MAX_STEPS, exponential_decay_lr, shard_range
    assert abs(exponential_decay_lr(0, LRS["means"], MEANS_LR_FINAL, MEANS_LR_MAX_STEPS) - LRS["means"]) < 1e-12
    assert abs(exponential_decay_lr(30000, LRS["means"], MEANS_LR_FINAL, MEANS_LR_MAX_STEPS) - MEANS_LR_FINAL) < 1e-12
    assert abs(exponential_decay_lr(99999, LRS["means"], MEANS_LR_FINAL, MEANS_LR_MAX_STEPS) - MEANS_LR_FINAL) < 1e-12
    mid = exponential_decay_lr(15000, LRS["means"], MEANS_LR_FINAL, MEANS_LR_MAX_STEPS)
    assert abs(mid - (LRS["means"] * MEANS_LR_FINAL) ** 0.5) < 1e-12
    assert mid == OO.exponential_decay_lr(15000, LRS["means"], MEANS_LR_FINAL, MEANS_LR_MAX_STEPS)
    for n in (0, 5, 48_000_003):
        for world in (1, 2, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo % 4 == 0 for lo, _ in r)


def _close(a, b, rel):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max()) <= rel * max(float(b.abs().max()), 1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 1023, 50_001])
def test_fused_adam_matches_torch_adam(built_lib, n):
    """12 steps, the reference's six groups; features_dc ++ features_rest stepped as ONE [N,16,3] segment with
    two learning rates.  Tolerance 2e-6 of the tensor's range (fp32 contraction differs between the CPU and
    the kernel; the update itself is operation-for-operation torch's)."""
    from freegaussian_b200 import _lib
    from freegaussian_b200.optim import GaussianAdam
    params, grads = _groups(n, n)
    ref, ref_state = OO.adam_reference(params, grads, LRS)
    dev = "cuda"
    sh = torch.cat([params["features_dc"][:, None], params["features_rest"]], 1).to(dev).contiguous()
    p = {k: params[k].to(dev).contiguous() for k in ("means", "opacities", "scales", "quats")}
    opt = GaussianAdam.for_reference_groups(p["means"], sh, p["opacities"], p["scales"], p["quats"])
    l0 = _lib.launch_count()
    for g in grads:
        gsh = torch.cat([g["features_dc"][:, None], g["features_rest"]], 1).to(dev).contiguous()
        opt.step({"means": g["means"].to(dev), "sh": gsh, "opacities": g["opacities"].to(dev),
                  "scales": g["scales"].to(dev), "quats": g["quats"].to(dev)})
    assert _lib.launch_count() - l0 == len(grads)  # one launch per step for all groups
    got = dict(p, features_dc=sh[:, 0], features_rest=sh[:, 1:])
    st = opt.state()
    got_m = {k: st[k][0] for k in p}
    got_v = {k: st[k][1] for k in p}
    got_m.update(features_dc=st["sh"][0][:, 0], features_rest=st["sh"][0][:, 1:])
    got_v.update(features_dc=st["sh"][1][:, 0], features_rest=st["sh"][1][:, 1:])
    for k in ref:
        assert _close(got[k], ref[k], 2e-6), k
        assert _close(got_m[k], ref_state[k][0], 2e-6), k
        assert _close(got_v[k], ref_state[k][1], 2e-6), k


@pytest.mark.gpu
def test_sharded_step_equals_full_step(built_lib):
    """Each of `world` ranks stepping its slice == one full step, bit for bit (the multi-GPU variant)."""
    from freegaussian_b200.optim import GaussianAdam
    params, grads = _groups(4099, 7)
    dev = "cuda"

    def make():
        sh = torch.cat([params["features_dc"][:, None], params["features_rest"]], 1).to(dev).contiguous()
        p = {k: params[k].to(dev).contiguous() for k in ("means", "opacities", "scales", "quats")}
        return GaussianAdam.for_reference_groups(p["means"], sh, p["opacities"], p["scales"], p["quats"])

    def dev_grads(g):
        gsh = torch.cat([g["features_dc"][:, None], g["features_rest"]], 1).to(dev).contiguous()
        return {"means": g["means"].to(dev), "sh": gsh, "opacities": g["opacities"].to(dev),
                "scales": g["scales"].to(dev), "quats": g["quats"].to(dev)}

    full, parts = make(), make()
    for g in grads[:4]:
        dg = dev_grads(g)
        full.step(dg)
        parts.t += 1
        for r in range(3):
            parts.t -= 1
            parts.step(dg, shard=(r, 3))
    for k in full.groups:
        assert torch.equal(full.groups[k].param, parts.groups[k].param), k
        assert torch.equal(full.groups[k].exp_avg_sq, parts.groups[k].exp_avg_sq), k


@pytest.mark.gpu
def test_rebind_after_refinement_keeps_the_step_count(built_lib):
    """After refine() every tensor is a new allocation: rebind() swaps them in and the bias correction carries on
    (the reference keeps `step` inside the surviving param_state, freegaussian_model.py:313-367)."""
    from freegaussian_b200.optim import GaussianAdam
    params, grads = _groups(300, 3)
    dev = "cuda"
    mk = lambda: GaussianAdam({"means": __import__("freegaussian_b200.optim", fromlist=["Group"]).Group(params["means"].to(dev).contiguous(), LRS["means"])})  # noqa: E731
    a, b = mk(), mk()
    for g in grads[:5]:
        a.step({"means": g["means"].to(dev)})
        b.step({"means": g["means"].to(dev)})
    keep = torch.arange(0, 300, 2, device=dev)  # a cull: half of the rows survive, moments travel with them
    st = b.state()["means"]
    b.rebind({"means": b.groups["means"].param[keep].contiguous()}, {"means": (st[0][keep].contiguous(), st[1][keep].contiguous())})
    assert b.t == 5
    for g in grads[5:9]:
        a.step({"means": g["means"].to(dev)})
        b.step({"means": g["means"].to(dev)[keep].contiguous()})
    assert torch.equal(a.groups["means"].param[keep], b.groups["means"].param)
    assert torch.equal(a.groups["means"].exp_avg[keep], b.groups["means"].exp_avg)
