"""The per-Gaussian math the sm_100a projection kernels call (csrc/splat_math.h), run on the
host through tests/host_harness and compared with the oracle (forward) and with
torch.autograd through the oracle (hand-derived VJPs).  CPU only."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import render as O
from util import small_scene, rel_err, grad_rel_err

ROOT = Path(__file__).resolve().parent.parent
_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def harness():
    out = ROOT / "build" / "host_harness.so"
    out.parent.mkdir(parents=True, exist_ok=True)
    src = ROOT / "tests" / "host_harness" / "harness.cpp"
    hdr = ROOT / "freegaussian_b200" / "csrc" / "splat_math.h"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        # -ffp-contract=off: keep host arithmetic un-fused so it is a fixed reference
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(out), str(src)])
    return C.CDLL(str(out))


def fp(t):
    return t.contiguous().numpy().ctypes.data_as(_f) if t is not None else None


def run_fwd(h, sc, W, H, sh_degree, with_next=True):
    Cn, N = sc.viewmats.shape[0], sc.means.shape[0]
    radii = np.zeros((Cn, N), np.int32)
    tiles = np.zeros((Cn, N), np.int32)
    m2d = np.zeros((Cn, N, 2), np.float32)
    dep = np.zeros((Cn, N), np.float32)
    con = np.zeros((Cn, N, 3), np.float32)
    comp = np.zeros((Cn, N), np.float32)
    rgb = np.zeros((Cn, N, 3), np.float32)
    flow = np.zeros((Cn, N, 2), np.float32)
    a = lambda x: x.ctypes.data_as(_f)
    h.h_project_fwd(Cn, N, fp(sc.means), fp(sc.quats), fp(sc.scales), fp(sc.viewmats), fp(sc.Ks), W, H,
                    C.c_float(0.3), C.c_float(0.01), C.c_float(1e10), C.c_float(0.0), 16, sh_degree, 16, fp(sc.sh),
                    fp(sc.means_next) if with_next else None, radii.ctypes.data_as(_i), a(m2d), a(dep), a(con),
                    a(comp), a(rgb), a(flow), tiles.ctypes.data_as(_i))
    return {k: torch.from_numpy(v) for k, v in dict(radii=radii, tiles=tiles, means2d=m2d, depths=dep, conics=con,
                                                    comps=comp, rgb=rgb, flow=flow).items()}


@pytest.mark.parametrize("sh_degree", [0, 1, 2, 3])
def test_projection_forward_matches_oracle(harness, sh_degree):
    W, H = 96, 64
    sc = small_scene(1200, W, H, views=3, seed=sh_degree)
    got = run_fwd(harness, sc, W, H, sh_degree)
    radii, m2d, dep, con, comp, _ = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, W, H)
    # radii may differ only where 3*sqrt(lambda) sits within rounding of an integer
    assert (got["radii"] != radii).float().mean() < 2e-3
    same = got["radii"] == radii
    assert same.float().mean() > 0.998
    vis = same & (radii > 0)
    assert vis.sum() > 300
    assert rel_err(got["means2d"][vis], m2d[vis]) < 1e-5
    assert rel_err(got["depths"][vis], dep[vis]) < 1e-6
    assert rel_err(got["comps"][vis], comp[vis]) < 1e-4
    # conics are 1/det-amplified: compare relative to each entry's own scale
    d = (got["conics"][vis] - con[vis]).abs() / (con[vis].abs().amax(-1, keepdim=True) + 1e-12)
    assert d.max() < 5e-4
    # colours
    campos = torch.inverse(sc.viewmats)[:, :3, 3]
    cols = O.spherical_harmonics(sh_degree, sc.means[None] - campos[:, None], sc.sh[None].expand(3, -1, -1, -1), radii > 0)
    cols = torch.clamp_min(cols + 0.5, 0.0)
    assert rel_err(got["rgb"][vis], cols[vis]) < 1e-5
    # flow features
    uv, z = O.project_points(sc.means_next, sc.viewmats, sc.Ks)
    ref_flow = torch.where(((radii > 0) & (z >= 0.01))[..., None], uv - m2d, torch.zeros(()))
    assert rel_err(got["flow"][vis], ref_flow[vis]) < 1e-4
    # tile counts
    x0, x1, y0, y1 = O.tile_rects(m2d, radii, 16, 6, 4)
    ref_tiles = ((x1 - x0) * (y1 - y0)).to(torch.int32)
    assert (got["tiles"][vis] != ref_tiles[vis]).float().mean() < 1e-3


@pytest.mark.parametrize("sh_degree", [0, 3])
def test_projection_vjp_matches_autograd(harness, sh_degree):
    W, H = 96, 64
    sc = small_scene(800, W, H, views=2, seed=7 + sh_degree)
    Cn, N = 2, 800
    params = [sc.means, sc.quats, sc.scales, sc.sh, sc.means_next]
    params = [p.clone().double().requires_grad_(True) for p in params]
    means, quats, scales, sh, means_next = params
    vm, Ks = sc.viewmats.double(), sc.Ks.double()
    radii, m2d, dep, con, comp, _ = O.fully_fused_projection(means, quats, scales, vm, Ks, W, H)
    vis = radii > 0
    campos = torch.inverse(vm)[:, :3, 3]
    cols = O.spherical_harmonics(sh_degree, means[None] - campos[:, None], sh[None].expand(Cn, -1, -1, -1), vis)
    cols = torch.clamp_min(cols + 0.5, 0.0)
    uv, z = O.project_points(means_next, vm, Ks)
    flow = torch.where((vis & (z >= 0.01))[..., None], uv - m2d, torch.zeros((), dtype=torch.double))
    g = torch.Generator().manual_seed(1)
    w = {k: torch.randn(v.shape, generator=g, dtype=torch.double) for k, v in
         dict(m2d=m2d, dep=dep, con=con, comp=comp, cols=cols, flow=flow).items()}
    loss = sum((w[k] * v).sum() for k, v in dict(m2d=m2d, dep=dep, con=con, comp=comp, cols=cols, flow=flow).items())
    loss.backward()

    f32 = lambda t: t.detach().float().contiguous()
    hr = run_fwd(harness, sc, W, H, sh_degree)
    assert (hr["radii"] == radii).all(), "pick another seed: borderline radius"
    v_means = np.zeros((N, 3), np.float32); v_quats = np.zeros((N, 4), np.float32)
    v_scales = np.zeros((N, 3), np.float32); v_sh = np.zeros((N, 16, 3), np.float32)
    v_next = np.zeros((N, 3), np.float32)
    a = lambda x: x.ctypes.data_as(_f)
    harness.h_project_bwd(Cn, N, fp(sc.means), fp(sc.quats), fp(sc.scales), fp(sc.viewmats), fp(sc.Ks), W, H,
                          C.c_float(0.3), C.c_float(0.01), C.c_float(1e10), C.c_float(0.0), sh_degree, 16, fp(sc.sh),
                          fp(sc.means_next), hr["radii"].numpy().ctypes.data_as(_i), fp(f32(w["m2d"])),
                          fp(f32(w["dep"])), fp(f32(w["con"])), fp(f32(w["comp"])), fp(f32(w["cols"])),
                          fp(f32(w["flow"])), a(v_means), a(v_quats), a(v_scales), a(v_sh), a(v_next))
    # float32 VJP vs float64 autograd: conic gradients are ill-conditioned, so compare on the
    # bulk (per-tensor max-normalised error) with the 1e-3 gradient tolerance of north_star
    for name, got, ref in [("means", v_means, means.grad), ("quats", v_quats, quats.grad),
                           ("scales", v_scales, scales.grad), ("sh", v_sh, sh.grad), ("next", v_next, means_next.grad)]:
        err = grad_rel_err(torch.from_numpy(got), ref)
        assert err < 1e-3, (name, err)


def test_flow_affine_and_vjp_match_autograd(harness):
    """Covariance flow term: A = B(t+1) B(t)^-1 - I and its hand-derived VJP."""
    g = torch.Generator().manual_seed(0)
    n = 500

    def rand_cov():
        L = torch.randn(n, 2, 2, generator=g, dtype=torch.double)
        S = L @ L.transpose(-1, -2) + 0.3 * torch.eye(2, dtype=torch.double)
        return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 1, 1]], -1)

    ct, cn = rand_cov().requires_grad_(True), rand_cov().requires_grad_(True)
    A = O.flow_affine_from_cov(ct, cn)
    w = torch.randn(n, 4, generator=g, dtype=torch.double)
    (A * w).sum().backward()
    f32 = lambda t: t.detach().float().contiguous()
    out = np.zeros((n, 4), np.float32)
    a = lambda x: x.ctypes.data_as(_f)
    harness.h_flow_affine(n, fp(f32(ct)), fp(f32(cn)), a(out))
    assert rel_err(torch.from_numpy(out), A) < 1e-5
    v_ct, v_cn = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    harness.h_flow_affine_vjp(n, fp(f32(ct)), fp(f32(cn)), fp(f32(w)), a(v_ct), a(v_cn))
    assert grad_rel_err(torch.from_numpy(v_ct), ct.grad) < 1e-4
    assert grad_rel_err(torch.from_numpy(v_cn), cn.grad) < 1e-4
