"""Stage-by-stage parity of the CUDA path against the oracle, through the C ABI (-m gpu)."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import render as O
from util import small_scene, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(built_lib):
    from freegaussian_b200 import _lib
    return _lib.lib()


def project_gpu(sc, W, H, sh_degree, with_next=True):
    from freegaussian_b200.rendering import _Project
    d = sc.to("cuda")
    cfg = dict(width=W, height=H, eps2d=0.3, near_plane=0.01, far_plane=1e10, radius_clip=0.0, tile_size=16,
               sh_degree=sh_degree, want_depth=True, antialiased=True)
    out = _Project.apply(d.means, d.quats, d.scales, d.sh, d.means_next if with_next else None, None, None, d.viewmats,
                         d.Ks, cfg)
    return [o.cpu() for o in out[:7]]


@pytest.mark.parametrize("sh_degree", [0, 1, 2, 3])
@pytest.mark.parametrize("recipe", ["trained_like", "init_like"])
def test_projection_matches_oracle(L, sh_degree, recipe):
    W, H = 200, 120  # partial right/bottom tiles
    sc = small_scene(5000, W, H, views=3, recipe=recipe, seed=sh_degree)
    radii, m2d, dep, con, comp, feat, tiles = project_gpu(sc, W, H, sh_degree)
    r_radii, r_m2d, r_dep, r_con, r_comp, _ = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, W, H)
    same = radii == r_radii
    assert same.float().mean() > 0.998  # ceil(3 sqrt(lambda)) may flip within rounding of an integer
    vis = same & (r_radii > 0)
    assert vis.sum() > 500
    assert rel_err(m2d[vis], r_m2d[vis]) < 1e-5
    assert rel_err(dep[vis], r_dep[vis]) < 1e-6
    assert rel_err(comp[vis], r_comp[vis]) < 1e-4
    d = (con[vis] - r_con[vis]).abs() / (r_con[vis].abs().amax(-1, keepdim=True) + 1e-12)
    assert d.max() < 5e-4
    campos = torch.inverse(sc.viewmats)[:, :3, 3]
    cols = O.spherical_harmonics(sh_degree, sc.means[None] - campos[:, None], sc.sh[None].expand(3, -1, -1, -1), r_radii > 0)
    cols = torch.clamp_min(cols + 0.5, 0.0)
    assert rel_err(feat[..., :3][vis], cols[vis]) < 1e-5
    assert torch.equal(feat[..., 3], dep)  # depth channel is the depth output, bit for bit
    uv, z = O.project_points(sc.means_next, sc.viewmats, sc.Ks)
    ref_flow = torch.where(((r_radii > 0) & (z >= 0.01))[..., None], uv - r_m2d, torch.zeros(()))
    assert rel_err(feat[..., 4:6][vis], ref_flow[vis]) < 1e-4
    # culled entries are zeroed
    assert (m2d[radii == 0] == 0).all() and (feat[radii == 0] == 0).all() and (tiles[radii == 0] == 0).all()
    # tile counts: bit-exact against the oracle rule applied to the kernel's own means2d/radii
    x0, x1, y0, y1 = O.tile_rects(m2d, radii, 16, math.ceil(W / 16), math.ceil(H / 16))
    assert torch.equal(tiles, ((x1 - x0) * (y1 - y0)).to(torch.int32))


def _isect(mode, *args):
    """isect_tiles in one of its list-building modes; "binned" places the coarse (splat, cell) pairs by rank when the call
    has few enough coarse cells, "binned_sorted" forces the emit + stable-sort variant (what larger calls use)."""
    from freegaussian_b200 import rendering
    if mode != "binned_sorted":
        return rendering.isect_tiles(*args, mode=mode)
    rendering.BIN_RANKED = False
    try:
        return rendering.isect_tiles(*args, mode="binned")
    finally:
        rendering.BIN_RANKED = True


def test_isect_sort_offsets_bit_exact(L):
    """Same projected tensors into both implementations: keys, order and tile ranges must be identical."""
    from freegaussian_b200.rendering import isect_tiles, isect_ids_from_tiles
    W, H = 200, 120
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    for seed, views, mul in [(0, 1, 1.0), (1, 3, 1.0), (2, 5, 4.0)]:
        sc = small_scene(4000, W, H, views=views, seed=seed, scale_mul=mul)
        radii, m2d, dep, con, comp, feat, tiles = project_gpu(sc, W, H, 0)
        tpg, r_ids, r_flat = O.isect_tiles(m2d, radii, dep, 16, tw, th)
        r_offs = O.isect_offset_encode(r_ids, views, tw, th)
        assert torch.equal(tpg, tiles)
        assert r_ids.numel() > 1000
        for mode in ("key64", "two_level", "binned", "binned_sorted"):
            ids, flat, offs, tk = _isect(mode, m2d.cuda(), radii.cuda(), dep.cuda(), tiles.cuda(), 16, tw, th)
            if ids is None and tk is not None:
                ids = isect_ids_from_tiles(tk, flat, dep.cuda(), tw, th)
            if ids is not None:
                assert torch.equal(ids.cpu(), r_ids), f"{mode}: sorted 64-bit keys differ"
            assert torch.equal(flat.cpu(), r_flat), f"{mode}: sort order differs"
            assert torch.equal(offs.cpu(), r_offs), f"{mode}: tile ranges differ"


@pytest.mark.parametrize("n,end_bit", [(1, 64), (5, 13), (4096, 40), (4097, 64), (100_003, 46), (3_000_000, 53)])
def test_radix_sort_u64_matches_stable_sort(L, n, end_bit):
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(0, 2**62, (n,), generator=g, dtype=torch.int64)
    keys &= (1 << end_bit) - 1 if end_bit < 63 else -1
    keys[::7] = keys[0]  # many duplicates: stability matters
    vals = torch.arange(n, dtype=torch.int32)
    ref_order = np.argsort(keys.numpy().astype(np.uint64), kind="stable")
    ka, va = keys.cuda(), vals.cuda()
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    ws = torch.empty(L.fg_radix_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    sel = ctypes.c_int(-1)
    rc = L.fg_radix_sort_pairs_u64_u32(n, ka.data_ptr(), va.data_ptr(), kb.data_ptr(), vb.data_ptr(), end_bit,
                                       ws.data_ptr(), ws.numel(), ctypes.byref(sel), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.fg_last_error()
    ko, vo = (kb, vb) if sel.value == 1 else (ka, va)
    assert np.array_equal(vo.cpu().numpy(), vals.numpy()[ref_order])
    assert np.array_equal(ko.cpu().numpy(), keys.numpy()[ref_order])


@pytest.mark.parametrize("n,end_bit", [(3, 32), (70_001, 32), (1_000_000, 17)])
def test_radix_sort_u32_matches_stable_sort(L, n, end_bit):
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(0, 2**end_bit, (n,), generator=g, dtype=torch.int64)
    vals = torch.arange(n, dtype=torch.int32)
    ref_order = np.argsort(keys.numpy(), kind="stable")
    ka, va = keys.to(torch.int32 if end_bit < 32 else torch.int64).cuda(), vals.cuda()
    ka = (keys & 0xFFFFFFFF).to(torch.int64).cuda().to(torch.int32) if end_bit == 32 else ka
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    ws = torch.empty(L.fg_radix_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    sel = ctypes.c_int(-1)
    rc = L.fg_radix_sort_pairs_u32_u32(n, ka.data_ptr(), va.data_ptr(), kb.data_ptr(), vb.data_ptr(), end_bit,
                                       ws.data_ptr(), ws.numel(), ctypes.byref(sel), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.fg_last_error()
    vo = vb if sel.value == 1 else va
    assert np.array_equal(vo.cpu().numpy(), vals.numpy()[ref_order])


def test_exclusive_scan(L):
    for n in [1, 255, 4096, 4097, 1_234_567]:
        g = torch.Generator().manual_seed(n)
        c = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
        cd = c.cuda()
        out = torch.empty_like(cd)
        tot = torch.zeros(1, dtype=torch.int64, device="cuda")
        ws = torch.empty(L.fg_scan_workspace_bytes(n), dtype=torch.uint8, device="cuda")
        rc = L.fg_exclusive_scan_i32(n, cd.data_ptr(), out.data_ptr(), tot.data_ptr(), ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream)
        assert rc == 0, L.fg_last_error()
        ref = torch.cumsum(c.long(), 0) - c.long()
        assert torch.equal(out.cpu().long(), ref)
        assert int(tot.item()) == int(c.long().sum())


@pytest.mark.parametrize("ch", [1, 3, 4, 6, 8])
def test_rasterize_fwd_bwd_matches_oracle(L, ch):
    """Same projected tensors + same sorted lists into both compositors."""
    from freegaussian_b200.rendering import isect_tiles, rasterize_to_pixels
    W, H = 100, 70
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    sc = small_scene(3000, W, H, views=2, seed=11 + ch)
    radii, m2d, dep, con, comp, feat, tiles = project_gpu(sc, W, H, 3)
    ids, flat, offs, _ = isect_tiles(m2d.cuda(), radii.cuda(), dep.cuda(), tiles.cuda(), 16, tw, th)
    g = torch.Generator().manual_seed(ch)
    cols = torch.rand(2, 3000, ch, generator=g)
    opac = sc.opacities[None].expand(2, -1).contiguous()
    bg = torch.rand(2, ch, generator=g)
    leaf = lambda t, dev: t.clone().to(dev).requires_grad_(True)
    # GPU
    gm, gc, gf, go, gb = leaf(m2d, "cuda"), leaf(con, "cuda"), leaf(cols, "cuda"), leaf(opac, "cuda"), leaf(bg, "cuda")
    gm2 = gm * 1.0  # non-leaf, like meta["means2d"]
    gm2.retain_grad()
    r, a = rasterize_to_pixels(gm2, gc, gf, go, W, H, 16, offs, flat, backgrounds=gb, absgrad=True)
    # oracle
    om, oc, of, oo, ob = leaf(m2d, "cpu"), leaf(con, "cpu"), leaf(cols, "cpu"), leaf(opac, "cpu"), leaf(bg, "cpu")
    rr, ra, last = O.rasterize_to_pixels(om, oc, of, oo, W, H, 16, offs.cpu(), flat.cpu(), backgrounds=ob)
    assert rel_err(r, rr) < 1e-4, rel_err(r, rr)
    assert rel_err(a, ra) < 1e-4
    wr = torch.randn(r.shape, generator=g)
    wa = torch.randn(a.shape, generator=g)
    ((r * wr.cuda()).sum() + (a * wa.cuda()).sum()).backward()
    ((rr * wr).sum() + (ra * wa).sum()).backward()
    from util import grad_rel_err
    for name, x, y in [("means2d", gm, om), ("conics", gc, oc), ("colors", gf, of), ("opac", go, oo), ("bg", gb, ob)]:
        e = grad_rel_err(x.grad, y.grad)
        assert e < 1e-3, (name, e)
    assert hasattr(gm2, "absgrad") and gm2.absgrad.shape == gm2.shape
    assert (gm2.absgrad >= gm2.grad.abs() - 1e-4 * gm2.absgrad.abs().max()).all()


def test_densify_stats_kernel_matches_reference_formula(L):
    """fg_densify_stats == the per-view arithmetic of freegaussian_model.py:376-392 (torch restatement)."""
    from freegaussian_b200.dist import DensificationStats
    g = torch.Generator().manual_seed(0)
    n = 10_001
    radii = torch.randint(0, 40, (3, n), generator=g, dtype=torch.int32)
    radii[radii < 15] = 0
    absgrad = torch.rand(3, n, 2, generator=g)
    a, b = DensificationStats(n, "cuda"), DensificationStats(n, "cpu")
    for _ in range(2):
        a.accumulate_local(radii.cuda(), absgrad.cuda(), 540, 960)
        b.accumulate_local(radii, absgrad, 540, 960)
        a.reduce(); b.reduce()
    assert torch.allclose(a.xys_grad_norm.cpu(), b.xys_grad_norm, rtol=1e-6, atol=1e-6)
    assert torch.equal(a.vis_counts.cpu(), b.vis_counts)
    assert torch.allclose(a.max_2Dsize.cpu(), b.max_2Dsize, rtol=1e-6)


@pytest.mark.parametrize("seed,C,W,H,n", [(0, 1, 50, 33, 3000), (1, 3, 333, 190, 20000), (2, 5, 1000, 37, 8000),
                                          (3, 2, 1920, 1080, 6000)])  # the last: 1020 coarse cells, the ranked path's limit
def test_tile_lists_adversarial(L, seed, C, W, H, n):
    """All three list builders against the oracle on hand-made splats: radii from 1 px to larger than the
    image, centres far outside, exact depth ties (tie-break = ascending c*N+n), ragged image sizes."""
    from freegaussian_b200.rendering import isect_tiles, isect_ids_from_tiles
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    g = torch.Generator().manual_seed(seed)
    m2d = (torch.rand(C, n, 2, generator=g) * 1.6 - 0.3) * torch.tensor([W, H])
    radii = torch.randint(1, 40, (C, n), generator=g, dtype=torch.int32)
    radii[:, ::17] = torch.randint(100, 2000, (C, len(range(0, n, 17))), generator=g, dtype=torch.int32)
    radii[torch.rand(C, n, generator=g) < 0.3] = 0
    dep = torch.rand(C, n, generator=g) * 10 + 0.1
    dep[:, ::5] = 1.25  # many exact ties
    dep[:, 1::7] = dep[:, 0:1]  # and ties with one particular value per camera
    m2d[radii == 0] = 0
    x0, x1, y0, y1 = O.tile_rects(m2d, radii, 16, tw, th)
    tiles = ((x1 - x0) * (y1 - y0)).to(torch.int32)
    tpg, r_ids, r_flat = O.isect_tiles(m2d, radii, dep, 16, tw, th)
    r_offs = O.isect_offset_encode(r_ids, C, tw, th)
    assert torch.equal(tpg, tiles) and r_ids.numel() > 0
    for mode in ("key64", "two_level", "binned", "binned_sorted"):
        ids, flat, offs, tk = _isect(mode, m2d.cuda(), radii.cuda(), dep.cuda(), tiles.cuda(), 16, tw, th)
        assert torch.equal(flat.cpu(), r_flat), f"{mode}: list order differs"
        assert torch.equal(offs.cpu(), r_offs), f"{mode}: tile ranges differ"
        if ids is None and tk is not None:
            ids = isect_ids_from_tiles(tk, flat, dep.cuda(), tw, th)
        if ids is not None:
            assert torch.equal(ids.cpu(), r_ids)
