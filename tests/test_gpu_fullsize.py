"""Render path at BASELINE.json's full size (cfg3: 1 M Gaussians, 1920x1080) through size-independent properties.

The oracle cannot composite a 1080p frame in test time, so at this size the checks are structural:
three independent tile-list builders must agree bit for bit, every tile list must be sorted by (depth, id) and
consistent with the offsets, compositing must be linear in the colours, packed and unpacked calls must agree, and
a window of the frame composited by the oracle from the Gaussians that can reach it must match the big frame.
"""
import pytest
import torch

from oracle import render as O
from util import rel_err

pytestmark = pytest.mark.gpu

N, W, H = 1_000_000, 1920, 1080


@pytest.fixture(scope="module")
def scene(built_lib):
    from freegaussian_b200.knn import k_nearest
    from freegaussian_b200.scenes import make_scene

    knn3 = lambda m: k_nearest(m.cuda(), 3)[0].cpu()  # noqa: E731  (bit-exact vs sklearn: tests/test_gpu_knn.py)
    return make_scene(N, W, H, n_views=1, recipe="trained_like", seed=0, knn3=knn3)


def _render(sc, **kw):
    from freegaussian_b200.rendering import rasterization

    d = sc.to("cuda")
    args = dict(packed=False, near_plane=0.01, far_plane=1e10, render_mode="RGB+ED", sh_degree=3, absgrad=True,
                rasterize_mode="classic", means_next=d.means_next)
    args.update(kw)
    colors = args.pop("colors", d.sh)
    with torch.no_grad():
        return rasterization(d.means, d.quats, d.scales, d.opacities, colors, d.viewmats, d.Ks, W, H, **args)


def test_three_tile_list_builders_agree_and_lists_are_sorted(scene):
    from freegaussian_b200 import rendering

    metas = {}
    old = rendering.SORT_MODE
    try:
        for mode in ("binned", "two_level", "key64"):
            rendering.SORT_MODE = mode
            metas[mode] = _render(scene)[2]
    finally:
        rendering.SORT_MODE = old
    ref = metas["key64"]  # the reference's literal 64-bit (camera | tile | depth) key sort
    for mode in ("binned", "two_level"):
        assert torch.equal(metas[mode]["flatten_ids"], ref["flatten_ids"]), mode
        assert torch.equal(metas[mode]["isect_offsets"], ref["isect_offsets"]), mode
        assert torch.equal(metas[mode]["radii"], ref["radii"]), mode
    m = ref
    ids, offs = m["flatten_ids"].long(), m["isect_offsets"].flatten().long()
    M = ids.numel()
    assert M == int(m["tiles_per_gauss"].sum()) and M > 10_000_000
    assert bool((offs[1:] >= offs[:-1]).all()) and int(offs[0]) == 0 and int(offs[-1]) <= M
    # inside a tile: depth bits non-decreasing, ties by ascending Gaussian id (the stable sort of the reference)
    depth_bits = m["depths"].flatten().view(torch.int32).long()[ids]
    same_tile = torch.ones(M - 1, dtype=torch.bool, device=ids.device)
    same_tile[offs[1:][(offs[1:] > 0) & (offs[1:] < M)] - 1] = False
    d0, d1 = depth_bits[:-1], depth_bits[1:]
    ok = (d1 > d0) | ((d1 == d0) & (ids[1:] > ids[:-1]))
    assert bool((ok | ~same_tile).all())
    # every listed Gaussian is visible
    assert bool((m["radii"].flatten()[ids] > 0).all())


def test_compositing_is_linear_in_the_colours(scene):
    g = torch.Generator().manual_seed(3)
    c1, c2 = torch.rand(N, 3, generator=g).cuda(), torch.rand(N, 3, generator=g).cuda()
    kw = dict(sh_degree=None, render_mode="RGB", means_next=None)
    r1, a1, _ = _render(scene, colors=c1, **kw)
    r2, a2, _ = _render(scene, colors=c2, **kw)
    r3, a3, _ = _render(scene, colors=0.25 * c1 + 2.0 * c2, **kw)
    assert torch.equal(a1, a2) and torch.equal(a1, a3)
    assert rel_err(r3, 0.25 * r1 + 2.0 * r2) < 1e-5
    assert float(a1.min()) >= 0.0 and float(a1.max()) <= 1.0 + 1e-6 and bool(torch.isfinite(r3).all())


def test_packed_call_matches_unpacked(scene):
    r, a, m = _render(scene)
    rp, ap, mp = _render(scene, packed=True)
    assert torch.equal(r, rp) and torch.equal(a, ap) and torch.equal(m["flow"], mp["flow"])
    vis = m["radii"].flatten() > 0
    assert torch.equal(mp["gaussian_ids"].long(), vis.nonzero().squeeze(1))
    assert torch.equal(mp["means2d"], m["means2d"].reshape(-1, 2)[vis])


def test_window_of_the_full_frame_matches_the_oracle(scene):
    """A 96x64 window of the 1080p frame, composited by the CPU oracle: projection, SH, tile lists and sort of the
    Gaussians that can reach the window at the FULL frame's camera (a shifted principal point would change gsplat's
    frustum clamp of the projection Jacobian), then the oracle's compositing over the window's 6x4 tiles only."""
    r, a, m = _render(scene)
    x0, y0, cw, ch = 912, 496, 96, 64  # tile aligned, near the centre
    m2d, rad = m["means2d"][0].cpu(), m["radii"][0].cpu().float()
    near = (rad > 0) & (m2d[:, 0] + rad > x0 - 2) & (m2d[:, 0] - rad < x0 + cw + 2) & (m2d[:, 1] + rad > y0 - 2) & (m2d[:, 1] - rad < y0 + ch + 2)
    idx = near.nonzero().squeeze(1)
    assert 100 < idx.numel() < 300_000
    means, quats, scales, opac, sh, mnext = (t[idx] for t in (scene.means, scene.quats, scene.scales, scene.opacities, scene.sh,
                                                              scene.means_next))
    vm, K = scene.viewmats, scene.Ks
    radii, means2d, depths, conics, _, _ = O.fully_fused_projection(means, quats, scales, vm, K, W, H, 0.3, 0.01, 1e10, 0.0)
    vis = radii > 0
    dirs = means[None] - torch.inverse(vm)[:, :3, 3][:, None]
    cols = torch.clamp_min(O.spherical_harmonics(3, dirs, sh[None], masks=vis) + 0.5, 0.0)
    uv_next, z_next = O.project_points(mnext, vm, K)
    flow2d = torch.where((vis & (z_next >= 0.01))[..., None], uv_next - means2d, torch.zeros(()))
    cols = torch.cat([cols, depths[..., None], flow2d], -1)  # rgb | depth | flow, as oracle.rasterization builds them
    tile_w, tile_h = (W + 15) // 16, (H + 15) // 16
    _, isect_ids, flatten_ids = O.isect_tiles(means2d, radii, depths, 16, tile_w, tile_h)
    offs = O.isect_offset_encode(isect_ids, 1, tile_w, tile_h).reshape(-1).tolist() + [flatten_ids.numel()]
    win_ids, win_offs = [], []
    for ty in range(y0 // 16, (y0 + ch) // 16):
        for tx in range(x0 // 16, (x0 + cw) // 16):
            t = ty * tile_w + tx
            win_offs.append(sum(x.numel() for x in win_ids))
            win_ids.append(flatten_ids[offs[t]:offs[t + 1]])
    win_ids = torch.cat(win_ids)
    win_offs = torch.tensor(win_offs, dtype=torch.int32).view(1, ch // 16, cw // 16)
    assert win_ids.numel() > 1000
    shifted = means2d - torch.tensor([float(x0), float(y0)])
    ro, ao, _ = O.rasterize_to_pixels(shifted, conics, cols, opac[None], cw, ch, 16, win_offs, win_ids)
    ro = torch.cat([ro[..., :3], ro[..., 3:4] / ao.clamp(min=1e-10), ro[..., 4:]], -1)  # "ED" normalisation
    crop = lambda t: t[:, y0:y0 + ch, x0:x0 + cw].cpu()  # noqa: E731
    assert rel_err(crop(r), ro[..., :4]) < 1e-4
    assert rel_err(crop(a), ao) < 1e-4
    assert rel_err(crop(m["flow"]), ro[..., 4:]) < 1e-4
