"""The oracle itself (CPU): golden fixtures, self-consistency in float64, analytic known answers,
finite differences, and the structural properties of keys / sort / tile ranges."""
import ast
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import render as O
from util import small_scene, rel_err

GOLD = Path(__file__).parent / "golden"


@pytest.mark.parametrize("name", ["rgbed_sh3", "rgb_sh0_aa"])
def test_render_golden(name):
    z = np.load(GOLD / f"render_{name}.npz")
    kw = ast.literal_eval(str(z["kwargs"]))
    t = lambda k: torch.from_numpy(z[k])
    r, a, m = O.rasterization(t("means"), t("quats"), t("scales"), t("opacities"), t("sh"), t("viewmats"), t("Ks"),
                              int(z["width"]), int(z["height"]), means_next=t("means_next"), **kw)
    assert np.array_equal(m["radii"].numpy(), z["radii"])
    assert np.array_equal(m["flatten_ids"].numpy(), z["flatten_ids"])
    assert np.array_equal(m["isect_ids"].numpy(), z["isect_ids"])
    assert np.array_equal(m["isect_offsets"].numpy(), z["isect_offsets"])
    assert np.array_equal(m["tiles_per_gauss"].numpy(), z["tiles_per_gauss"])
    assert rel_err(r, t("render")) < 1e-6 and rel_err(a, t("alpha")) < 1e-6 and rel_err(m["flow"], t("flow")) < 1e-6


def test_float32_vs_float64():
    W, H = 64, 48
    sc = small_scene(500, W, H, views=1, seed=31)
    a32 = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, sc.Ks, W, H, sh_degree=3,
                          render_mode="RGB+ED", means_next=sc.means_next)
    d = lambda x: x.double()
    a64 = O.rasterization(d(sc.means), d(sc.quats), d(sc.scales), d(sc.opacities), d(sc.sh), d(sc.viewmats), d(sc.Ks),
                          W, H, sh_degree=3, render_mode="RGB+ED", means_next=d(sc.means_next))
    assert torch.equal(a32[2]["radii"], a64[2]["radii"])
    assert rel_err(a32[0], a64[0]) < 1e-4 and rel_err(a32[1], a64[1]) < 1e-4
    assert rel_err(a32[2]["flow"], a64[2]["flow"]) < 1e-4


def test_single_gaussian_known_answer():
    """One isotropic Gaussian at the optical axis: alpha(p) = min(.999, o * exp(-|p-mu|^2 / (2 s2d))) with
    s2d = (f s / z)^2 + 0.3 (EWA blur), colour = SH DC * 0.2820948 + 0.5 (Appendix A.2/A.3/A.6)."""
    W = H = 32
    f, z, s, o = 40.0, 4.0, 0.2, 0.8
    means = torch.tensor([[0.0, 0.0, z]])
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    scales = torch.full((1, 3), s)
    sh = torch.zeros(1, 16, 3)
    sh[0, 0] = torch.tensor([1.0, -0.5, 0.25])
    vm = torch.eye(4)[None]
    K = torch.tensor([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]]])
    r, a, m = O.rasterization(means, quats, scales, torch.tensor([o]), sh, vm, K, W, H, sh_degree=0, render_mode="RGB+ED")
    s2d = (f * s / z) ** 2 + 0.3
    ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
    d2 = (xs - W / 2) ** 2 + (ys - H / 2) ** 2
    alpha = torch.clamp(o * torch.exp(-0.5 * d2 / s2d), max=0.999)
    alpha = torch.where(alpha >= 1 / 255, alpha, torch.zeros(()))
    radius = math.ceil(3 * math.sqrt(s2d))
    assert int(m["radii"][0, 0]) == radius
    assert rel_err(a[0, ..., 0], alpha) < 1e-6
    col = torch.clamp_min(sh[0, 0] * 0.2820947917738781 + 0.5, 0)
    assert rel_err(r[0, ..., :3], alpha[..., None] * col) < 1e-6
    assert rel_err(r[0, ..., 3][alpha > 0], torch.full_like(alpha, z)[alpha > 0]) < 1e-5  # ED = depth where covered (fp32 divide)


def test_gradients_match_finite_differences():
    W, H = 24, 16
    sc = small_scene(40, W, H, views=1, seed=5, scale_mul=1.5)
    d = lambda x: x.double()
    base = [d(sc.means), d(sc.quats), d(sc.scales), d(sc.opacities), d(sc.sh), d(sc.means_next)]
    g = torch.Generator().manual_seed(0)
    wr = torch.randn(1, H, W, 4, generator=g, dtype=torch.double)
    wf = torch.randn(1, H, W, 2, generator=g, dtype=torch.double)

    def loss(ps):
        r, a, m = O.rasterization(ps[0], ps[1], ps[2], ps[3], ps[4], d(sc.viewmats), d(sc.Ks), W, H, sh_degree=3,
                                  render_mode="RGB+ED", means_next=ps[5])
        return (r * wr).sum() + (m["flow"] * wf).sum() + a.sum()

    ps = [p.clone().requires_grad_(True) for p in base]
    loss(ps).backward()
    rng = np.random.default_rng(0)
    for pi in range(6):
        for _ in range(3):
            idx = tuple(int(rng.integers(0, s)) for s in base[pi].shape)
            eps = 1e-6
            plus = [p.clone() for p in base]; plus[pi][idx] += eps
            minus = [p.clone() for p in base]; minus[pi][idx] -= eps
            fd = float(loss(plus) - loss(minus)) / (2 * eps)
            an = float(ps[pi].grad[idx])
            assert abs(fd - an) <= 1e-4 * max(1.0, abs(an), abs(fd)), (pi, idx, fd, an)


def test_keys_sort_and_offsets_properties():
    W, H = 120, 72
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    sc = small_scene(1500, W, H, views=3, seed=8)
    radii, m2d, dep, con, comp, _ = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, W, H)
    tpg, ids, flat = O.isect_tiles(m2d, radii, dep, 16, tw, th)
    offs = O.isect_offset_encode(ids, 3, tw, th)
    ids_n, flat_n = ids.numpy(), flat.numpy()
    assert (np.diff(ids_n) >= 0).all()  # sortedness
    same = np.diff(ids_n) == 0
    assert (np.diff(flat_n)[same] > 0).all()  # stability: ties keep ascending c*N+n
    tile_bits = int(math.floor(math.log2(tw * th))) + 1
    cam = ids_n >> (32 + tile_bits)
    assert np.array_equal(cam, flat_n // 1500)  # camera field agrees with the value
    depth_bits = (ids_n & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    assert np.array_equal(depth_bits, dep.reshape(-1).numpy()[flat_n])
    assert int(tpg.sum()) == ids.numel()
    o = offs.reshape(-1).numpy()
    assert (np.diff(o) >= 0).all() and o[0] == 0 and o[-1] <= ids.numel()
    # every (c,n) appears exactly tiles_per_gauss times
    assert np.array_equal(np.bincount(flat_n, minlength=3 * 1500), tpg.reshape(-1).numpy())


def test_flow_equals_extra_colour_channels():
    W, H = 64, 48
    sc = small_scene(400, W, H, views=1, seed=12)
    r, a, m = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, sc.Ks, W, H, sh_degree=3,
                              means_next=sc.means_next)
    uv, z = O.project_points(sc.means_next, sc.viewmats, sc.Ks)
    f = torch.where(((m["radii"] > 0) & (z >= 0.01))[..., None], uv - m["means2d"], torch.zeros(()))
    r2, _, _ = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, f[0], sc.viewmats, sc.Ks, W, H, sh_degree=None)
    assert rel_err(m["flow"], r2) < 1e-6


def test_packed_meta_layout():
    W, H = 48, 32
    sc = small_scene(300, W, H, views=2, seed=2)
    r0, a0, m0 = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, sc.Ks, W, H, sh_degree=3,
                                 render_mode="ED", packed=False)
    r1, a1, m1 = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, sc.Ks, W, H, sh_degree=3,
                                 render_mode="ED", packed=True)
    assert torch.equal(r0, r1)
    nnz = int((m0["radii"] > 0).sum())
    assert m1["means2d"].shape == (nnz, 2) and m1["gaussian_ids"].shape == (nnz,)
    assert (np.diff((m1["camera_ids"] * 300 + m1["gaussian_ids"]).numpy()) > 0).all()


def test_compositing_matches_a_scalar_pixel_loop():
    """Second, independent statement of Appendix A.4-A.6: per tile the Gaussians whose tile rectangle covers it, sorted by
    (float32 depth bits, index) with Python's sort; per pixel a plain loop (skip sigma < 0 or alpha < 1/255, alpha capped at
    0.999, stop BEFORE the Gaussian that would take T to <= 1e-4).  Cross-checks the oracle's vectorised emission, key sort,
    offsets, cumulative-product compositing and `last_ids` against 30 lines of scalar code, in float64."""
    W, H, ts = 56, 40, 16
    sc = small_scene(260, W, H, views=1, seed=11)
    dbl = lambda t: t.double()  # noqa: E731
    r, a, m = O.rasterization(dbl(sc.means), dbl(sc.quats), dbl(sc.scales), dbl(sc.opacities), dbl(sc.sh), dbl(sc.viewmats), dbl(sc.Ks),
                              W, H, render_mode="RGB+ED", sh_degree=3, means_next=dbl(sc.means_next))
    radii, m2, conics, depths, opac, cols = (m[k][0].numpy() for k in ("radii", "means2d", "conics", "depths", "opacities", "colors"))
    tile_w, tile_h = math.ceil(W / ts), math.ceil(H / ts)
    m2f, rf = m2.astype(np.float32), radii.astype(np.float32)
    depth_key = depths.astype(np.float32).view(np.int32).astype(np.int64) & 0xFFFFFFFF
    out = np.zeros((H, W, cols.shape[1]))
    alpha_img = np.zeros((H, W))
    last = np.zeros((H, W), np.int64)
    pos = 0  # running position in the global sorted list = what last_ids indexes
    for ty in range(tile_h):
        for tx in range(tile_w):
            lst = []
            for g in np.nonzero(radii > 0)[0]:
                x0 = min(max(math.floor(m2f[g, 0] / np.float32(ts) - rf[g] / np.float32(ts)), 0), tile_w)
                x1 = min(max(math.ceil(m2f[g, 0] / np.float32(ts) + rf[g] / np.float32(ts)), 0), tile_w)
                y0 = min(max(math.floor(m2f[g, 1] / np.float32(ts) - rf[g] / np.float32(ts)), 0), tile_h)
                y1 = min(max(math.ceil(m2f[g, 1] / np.float32(ts) + rf[g] / np.float32(ts)), 0), tile_h)
                if x0 <= tx < x1 and y0 <= ty < y1:
                    lst.append((int(depth_key[g]), int(g)))
            lst.sort()
            for py in range(ty * ts, min((ty + 1) * ts, H)):
                for px in range(tx * ts, min((tx + 1) * ts, W)):
                    T, acc = 1.0, np.zeros(cols.shape[1])
                    for k, (_, g) in enumerate(lst):
                        dx, dy = m2[g, 0] - (px + 0.5), m2[g, 1] - (py + 0.5)
                        sigma = 0.5 * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
                        alpha = min(0.999, opac[g] * math.exp(-sigma))
                        if sigma < 0 or alpha < 1.0 / 255.0:
                            continue
                        if T * (1 - alpha) <= 1e-4:
                            break
                        acc += alpha * T * cols[g]
                        T *= 1 - alpha
                        last[py, px] = pos + k
                    out[py, px], alpha_img[py, px] = acc, 1 - T
            pos += len(lst)
    assert pos == m["flatten_ids"].numel()
    want = np.concatenate([out[..., :3], out[..., 3:4] / np.maximum(alpha_img[..., None], 1e-10)], -1)  # "ED"
    assert np.abs(r[0].numpy() - want).max() < 1e-9
    assert np.abs(a[0, ..., 0].numpy() - alpha_img).max() < 1e-12
    assert np.abs(m["flow"][0].numpy() - out[..., 4:6]).max() < 1e-9
    assert np.array_equal(m["last_ids"][0].numpy(), last)


def test_projection_matches_a_scalar_restatement():
    """Appendix A.2 per Gaussian with plain Python floats (float64), from the spec, against the oracle's vectorised
    `fully_fused_projection` run in float64: radii exact, means2d / depths / conics / compensation to 1e-10."""
    W, H = 72, 40
    sc = small_scene(400, W, H, views=2, seed=23)
    mu, q, s = sc.means.double().numpy(), sc.quats.double().numpy(), sc.scales.double().numpy()
    vms, Ks = sc.viewmats.double().numpy(), sc.Ks.double().numpy()
    radii, means2d, depths, conics, comp, _ = O.fully_fused_projection(
        sc.means.double(), sc.quats.double(), sc.scales.double(), sc.viewmats.double(), sc.Ks.double(), W, H, 0.3, 0.01, 1e10, 0.0)
    n_vis = 0
    for c in range(vms.shape[0]):
        Rcw, tcw = vms[c, :3, :3], vms[c, :3, 3]
        fx, fy, cx, cy = Ks[c, 0, 0], Ks[c, 1, 1], Ks[c, 0, 2], Ks[c, 1, 2]
        for n in range(mu.shape[0]):
            w_, x_, y_, z_ = q[n] / np.linalg.norm(q[n])
            R = np.array([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - w_ * z_), 2 * (x_ * z_ + w_ * y_)],
                          [2 * (x_ * y_ + w_ * z_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - w_ * x_)],
                          [2 * (x_ * z_ - w_ * y_), 2 * (y_ * z_ + w_ * x_), 1 - 2 * (x_ * x_ + y_ * y_)]])
            M = R * s[n][None, :]
            cov_c = Rcw @ (M @ M.T) @ Rcw.T
            x, y, z = Rcw @ mu[n] + tcw
            r_want = 0
            if 0.01 <= z <= 1e10:
                tan_x, tan_y = 0.5 * W / fx, 0.5 * H / fy
                tx = z * min(max(x / z, -(cx / fx + 0.3 * tan_x)), (W - cx) / fx + 0.3 * tan_x)
                ty = z * min(max(y / z, -(cy / fy + 0.3 * tan_y)), (H - cy) / fy + 0.3 * tan_y)
                J = np.array([[fx / z, 0.0, -fx * tx / (z * z)], [0.0, fy / z, -fy * ty / (z * z)]])
                c2 = J @ cov_c @ J.T
                det_orig = c2[0, 0] * c2[1, 1] - c2[0, 1] ** 2
                a_, b_, c_ = c2[0, 0] + 0.3, c2[0, 1], c2[1, 1] + 0.3
                det = a_ * c_ - b_ * b_
                u, v = fx * x / z + cx, fy * y / z + cy
                if det > 0:
                    mid = 0.5 * (a_ + c_)
                    rad = math.ceil(3.0 * math.sqrt(mid + math.sqrt(max(0.01, mid * mid - det))))
                    if not (u + rad <= 0 or u - rad >= W or v + rad <= 0 or v - rad >= H):
                        r_want = rad
            assert int(radii[c, n]) == r_want, (c, n)
            if r_want:
                n_vis += 1
                assert abs(means2d[c, n, 0] - u) < 1e-9 and abs(means2d[c, n, 1] - v) < 1e-9 and abs(depths[c, n] - z) < 1e-12
                want_conic = np.array([c_ / det, -b_ / det, a_ / det])
                assert np.abs(conics[c, n].numpy() - want_conic).max() < 1e-10 * max(1.0, np.abs(want_conic).max())
                assert abs(comp[c, n] - math.sqrt(max(0.0, det_orig / det))) < 1e-10
    assert n_vis > 200


def test_sh_basis_is_the_real_spherical_harmonics():
    """The 16 basis functions of Appendix A.3 against their mathematical definition: the real spherical harmonics with the
    Condon-Shortley phase, Y_l0, sqrt(2) Re Y_l^m (m > 0), sqrt(2) Im Y_l^|m| (m < 0), from scipy's complex Y_l^m."""
    import scipy.special as sp

    rng = np.random.default_rng(0)
    d = rng.standard_normal((64, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    theta, phi = np.arccos(d[:, 2]), np.arctan2(d[:, 1], d[:, 0])
    basis = O.eval_sh_bases(3, torch.from_numpy(d)).numpy()
    k = 0
    for l in range(4):  # noqa: E741
        for m in range(-l, l + 1):
            Y = sp.sph_harm_y(l, abs(m), theta, phi) if hasattr(sp, "sph_harm_y") else sp.sph_harm(abs(m), l, phi, theta)
            want = Y.real if m == 0 else np.sqrt(2) * (Y.real if m > 0 else Y.imag)
            assert np.abs(basis[:, k] - want).max() < 1e-14, (l, m)
            k += 1
    assert k == 16 == O.num_sh_bases(3)


def test_absgrad_tap_equals_a_pixel_by_pixel_backward():
    """``meta["absgrad"]`` (gsplat's ``means2d.absgrad``, consumed at ``freegaussian_model.py:377``) is the sum over pixels of
    |d loss_p / d means2d|: check the autograd tap against one backward per pixel on a tiny frame."""
    W, H = 12, 9
    sc = small_scene(40, W, H, views=1, seed=23, scale_mul=1.5)
    g = torch.Generator().manual_seed(2)
    w = torch.randn(1, H, W, 3, generator=g).double()
    d = lambda x: x.double()
    args = (d(sc.means).requires_grad_(True), d(sc.quats), d(sc.scales), d(sc.opacities), d(sc.sh), d(sc.viewmats),
            d(sc.Ks), W, H)
    r, a, m = O.rasterization(*args, sh_degree=3, absgrad=True)
    assert int((m["radii"] > 0).sum()) > 10
    m["means2d"].retain_grad()
    (r * w).sum().backward()
    want = torch.zeros_like(m["absgrad"])
    for y in range(H):
        for x in range(W):
            r2, _, m2 = O.rasterization(*args, sh_degree=3)
            m2["means2d"].retain_grad()
            (r2[0, y, x] * w[0, y, x]).sum().backward()
            want += m2["means2d"].grad.abs()
    assert float(want.max()) > 0
    assert rel_err(m["absgrad"], want) < 1e-12
    # and the tap leaves the ordinary gradient untouched
    r3, _, m3 = O.rasterization(*args, sh_degree=3)
    m3["means2d"].retain_grad()
    (r3 * w).sum().backward()
    assert torch.allclose(m["means2d"].grad, m3["means2d"].grad, rtol=0, atol=1e-14)
