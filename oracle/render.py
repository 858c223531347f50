"""Pure-torch CPU restatement of the splat-render path FreeGaussian calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED.  The arithmetic of this path lives in the un-vendored
dependency ``gsplat`` (``/root/reference/pyproject.toml:9``: ``gsplat >= 1.0.0``,
no upper pin, no lock file; effective range 1.0.0-1.4.x because
``freegaussian/freegaussian_model.py:15,21`` import ``gsplat.cuda_legacy`` and
``:376,872`` need 2-D ``radii``).  gsplat is absent from ``/root/reference`` and
is not installable here, and the reference ships no tests, golden vectors or
fixtures for this path.  This file therefore restates
gsplat 1.4.0's published algorithm and is anchored on the
reference's own call sites:

* ``freegaussian/freegaussian_model.py:847-868`` and
  ``freegaussian/freegaussian_control_model.py:158-179`` -- the keyword surface
  (``tile_size=16, packed=False, near_plane=0.01, far_plane=1e10,
  render_mode in {"RGB","RGB+ED"}, sh_degree in {None,0..3}, sparse_grad=False,
  absgrad=True, rasterize_mode in {"classic","antialiased"}``);
* ``preprocess/knn_gaussian.py:93-113`` etc. -- ``packed=True``, ``"ED"``;
* ``freegaussian/utils.py:162-179`` -- camera convention (world->camera, OpenCV axes);
* ``freegaussian/utils.py:232-245`` -- SH DC constant;
* ``docs/index.html:286-299`` -- Corollary 1, the only definition of the rendered
  Gaussian flow: ``u = sum_i T_i alpha_i (mu_{i,t+1} - mu_{i,t})``.

Everything is vectorised torch on CPU; the backward pass is ``torch.autograd``
through these functions, so it doubles as the gradient oracle.  ``dtype`` can be
float64 for finite-difference checks.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

ALPHA_MIN = 1.0 / 255.0  # Appendix A.6: skip if alpha < 1/255
ALPHA_MAX = 0.999  # Appendix A.6: alpha = min(0.999, ...)
T_STOP = 1e-4  # Appendix A.6: stop if T(1-alpha) <= 1e-4


# --------------------------------------------------------------------------- A.2
def quat_to_rotmat(quats: Tensor) -> Tensor:
    """(w,x,y,z) -> rotation matrix, normalising inside (Appendix A.1/A.2).

    Stands in for ``gsplat.cuda_legacy._torch_impl.quat_to_rotmat`` used at
    ``freegaussian_model.py:15,535``.
    """
    q = quats / quats.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=-1,
    )
    return R.reshape(quats.shape[:-1] + (3, 3))


def quat_scale_to_covar(quats: Tensor, scales: Tensor) -> Tensor:
    """Sigma = M M^T with M = R(q) diag(s)  (Appendix A.2)."""
    M = quat_to_rotmat(quats) * scales[..., None, :]
    return M @ M.transpose(-1, -2)


def project_points(means: Tensor, viewmats: Tensor, Ks: Tensor) -> Tuple[Tensor, Tensor]:
    """World points -> (pixel coordinates [C,N,2], camera depth [C,N]).  No clamping."""
    R = viewmats[:, :3, :3]
    t = viewmats[:, :3, 3]
    pc = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]
    x, y, z = pc.unbind(-1)
    fx, fy, cx, cy = Ks[:, 0, 0], Ks[:, 1, 1], Ks[:, 0, 2], Ks[:, 1, 2]
    u = fx[:, None] * x / z + cx[:, None]
    v = fy[:, None] * y / z + cy[:, None]
    return torch.stack([u, v], -1), z


def project_cov2d(means, quats, scales, viewmats, Ks, width, height):
    """Un-blurred 2-D covariance [C,N,2,2], projected mean [C,N,2] and camera depth [C,N]; no culling."""
    C, N = viewmats.shape[0], means.shape[0]
    covars = quat_scale_to_covar(quats, scales)  # [N,3,3]
    R = viewmats[:, :3, :3]
    t = viewmats[:, :3, 3]
    pc = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]  # [C,N,3]
    cov_c = torch.einsum("cij,njk,clk->cnil", R, covars, R)  # R Sigma R^T
    x, y, z = pc.unbind(-1)
    fx, fy = Ks[:, 0, 0][:, None], Ks[:, 1, 1][:, None]
    cx, cy = Ks[:, 0, 2][:, None], Ks[:, 1, 2][:, None]
    tan_fovx = 0.5 * width / fx
    tan_fovy = 0.5 * height / fy
    lim_x_pos = (width - cx) / fx + 0.3 * tan_fovx
    lim_x_neg = cx / fx + 0.3 * tan_fovx
    lim_y_pos = (height - cy) / fy + 0.3 * tan_fovy
    lim_y_neg = cy / fy + 0.3 * tan_fovy
    rz = 1.0 / z
    rz2 = rz * rz
    tx = z * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, x * rz))
    ty = z * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, y * rz))
    O = torch.zeros_like(z)
    J = torch.stack([fx * rz, O, -fx * tx * rz2, O, fy * rz, -fy * ty * rz2], -1).reshape(C, N, 2, 3)
    cov2d = J @ cov_c @ J.transpose(-1, -2)  # [C,N,2,2]
    means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], -1)
    return cov2d, means2d, z


def fully_fused_projection(
    means: Tensor,  # [N,3]
    quats: Tensor,  # [N,4]
    scales: Tensor,  # [N,3]
    viewmats: Tensor,  # [C,4,4]
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """EWA projection with culling (Appendix A.2).

    Returns radii [C,N] int32 (0 = culled), means2d [C,N,2], depths [C,N],
    conics [C,N,3], compensations [C,N], cov2d (blurred) [C,N,3] = (a,b,c).
    Values at culled entries are zeroed.
    """
    cov2d, means2d, z = project_cov2d(means, quats, scales, viewmats, Ks, width, height)

    a0, b0, c0 = cov2d[..., 0, 0], cov2d[..., 0, 1], cov2d[..., 1, 1]
    det_orig = a0 * c0 - b0 * b0
    a, b, c = a0 + eps2d, b0, c0 + eps2d
    det = a * c - b * b
    compensation = torch.sqrt(torch.clamp(det_orig / det, min=0.0))
    conics = torch.stack([c / det, -b / det, a / det], -1)

    with torch.no_grad():
        mid = 0.5 * (a + c)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.01))
        radius = torch.ceil(3.0 * torch.sqrt(lam))
        valid = (z >= near_plane) & (z <= far_plane) & (det > 0)
        valid &= radius > radius_clip
        valid &= (means2d[..., 0] + radius > 0) & (means2d[..., 0] - radius < width)
        valid &= (means2d[..., 1] + radius > 0) & (means2d[..., 1] - radius < height)
        radius = torch.nan_to_num(radius, nan=0.0, posinf=0.0, neginf=0.0)
        radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)

    zero = torch.zeros((), dtype=means.dtype)
    means2d = torch.where(valid[..., None], means2d, zero)
    depths = torch.where(valid, z, zero)
    conics = torch.where(valid[..., None], conics, zero)
    compensation = torch.where(valid, compensation, zero)
    cov2d_blur = torch.where(valid[..., None], torch.stack([a, b, c], -1), zero)
    return radii, means2d, depths, conics, compensation, cov2d_blur


# --------------------------------------------------------------------------- A.3
def num_sh_bases(degree: int) -> int:
    """(degree+1)^2 -- stands in for ``gsplat.cuda_legacy._wrapper.num_sh_bases``
    (``freegaussian_model.py:21,165``)."""
    return (degree + 1) ** 2


def eval_sh_bases(degree: int, dirs: Tensor) -> Tensor:
    """Real SH basis values [..., (degree+1)^2] for *normalised* dirs (Appendix A.3)."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, 0.2820947917738781)]
    if degree >= 1:
        out += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if degree >= 2:
        z2 = z * z
        fTmp0B = -1.092548430592079 * z
        fC1 = x * x - y * y
        fS1 = 2.0 * x * y
        out += [
            0.5462742152960395 * fS1,
            fTmp0B * y,
            0.9461746957575601 * z2 - 0.3153915652525201,
            fTmp0B * x,
            0.5462742152960395 * fC1,
        ]
    if degree >= 3:
        fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
        fTmp1B = 1.445305721320277 * z
        fC2 = x * fC1 - y * fS1
        fS2 = x * fS1 + y * fC1
        out += [
            -0.5900435899266435 * fS2,
            fTmp1B * fS1,
            fTmp0C * y,
            z * (1.865881662950577 * z2 - 1.119528997770346),
            fTmp0C * x,
            fTmp1B * fC1,
            -0.5900435899266435 * fC2,
        ]
    return torch.stack(out, -1)


def spherical_harmonics(degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """dirs [...,3] (un-normalised), coeffs [...,K,3] -> colours [...,3]; zero where masked out."""
    K = num_sh_bases(degree)
    d = dirs / dirs.norm(dim=-1, keepdim=True)
    basis = eval_sh_bases(degree, d)  # [...,K]
    col = (basis[..., None] * coeffs[..., :K, :]).sum(-2)
    if masks is not None:
        col = torch.where(masks[..., None], col, torch.zeros((), dtype=col.dtype))
    return col


# --------------------------------------------------------------------------- A.4 / A.5
def tile_rects(means2d: Tensor, radii: Tensor, tile_size: int, tile_w: int, tile_h: int):
    """Inclusive-min / exclusive-max tile rectangle per (c,n) (Appendix A.4).  float32 arithmetic."""
    m = means2d.detach().to(torch.float32)
    r = radii.to(torch.float32)
    tx = m[..., 0] / tile_size
    ty = m[..., 1] / tile_size
    tr = r / tile_size
    x0 = torch.clamp(torch.floor(tx - tr), 0, tile_w).to(torch.int64)
    x1 = torch.clamp(torch.ceil(tx + tr), 0, tile_w).to(torch.int64)
    y0 = torch.clamp(torch.floor(ty - tr), 0, tile_h).to(torch.int64)
    y1 = torch.clamp(torch.ceil(ty + tr), 0, tile_h).to(torch.int64)
    vis = radii > 0
    z = torch.zeros_like(x0)
    return torch.where(vis, x0, z), torch.where(vis, x1, z), torch.where(vis, y0, z), torch.where(vis, y1, z)


def isect_tiles(
    means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_w: int, tile_h: int, sort: bool = True
) -> Tuple[Tensor, Tensor, Tensor]:
    """Emit and (optionally) stably sort (key64, flatten_id) pairs (Appendix A.4/A.5).

    key = cam << (32+tile_bits) | tile << 32 | float32 bits of depth.
    Returns tiles_per_gauss [C,N] int32, isect_ids [M] int64, flatten_ids [M] int32.
    """
    C, N = radii.shape
    x0, x1, y0, y1 = tile_rects(means2d, radii, tile_size, tile_w, tile_h)
    tpg = ((x1 - x0) * (y1 - y0)).to(torch.int32)
    tile_bits = int(math.floor(math.log2(tile_w * tile_h))) + 1
    depth_bits = depths.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    x0n, x1n, y0n, y1n = (v.reshape(-1).numpy() for v in (x0, x1, y0, y1))
    cnt = tpg.reshape(-1).numpy().astype(np.int64)
    M = int(cnt.sum())
    flat = np.repeat(np.arange(C * N, dtype=np.int64), cnt)
    start = np.cumsum(cnt) - cnt
    local = np.arange(M, dtype=np.int64) - np.repeat(start, cnt)
    wid = np.repeat((x1n - x0n), cnt)
    wid = np.maximum(wid, 1)
    ti = np.repeat(y0n, cnt) + local // wid  # rows outer
    tj = np.repeat(x0n, cnt) + local % wid  # columns inner
    cam = flat // N
    keys = (cam << (32 + tile_bits)) | ((ti * tile_w + tj) << 32) | depth_bits.reshape(-1).numpy()[flat]
    if sort:
        order = np.argsort(keys, kind="stable")
        keys, flat = keys[order], flat[order]
    return tpg, torch.from_numpy(keys.astype(np.int64)), torch.from_numpy(flat.astype(np.int32))


def isect_offset_encode(isect_ids: Tensor, C: int, tile_w: int, tile_h: int) -> Tensor:
    """offsets[c,i,j] = lower bound of (c, i*tile_w+j) in the sorted keys (Appendix A.5)."""
    tile_bits = int(math.floor(math.log2(tile_w * tile_h))) + 1
    ids = isect_ids.numpy() >> 32
    cam = ids >> tile_bits
    tile = ids & ((1 << tile_bits) - 1)
    lin = cam * (tile_w * tile_h) + tile
    q = np.arange(C * tile_w * tile_h, dtype=np.int64)
    off = np.searchsorted(lin, q, side="left").astype(np.int32)
    return torch.from_numpy(off).reshape(C, tile_h, tile_w)


# --------------------------------------------------------------------------- A.6
class _AbsGradTap(torch.autograd.Function):
    """Identity on a per-pixel copy ``[G,P]`` of one coordinate of the tile's 2-D means.  Its backward sees
    d(loss)/d(mu) pixel by pixel and adds ``sum_p |.|`` into ``sink[g, col]`` -- gsplat's ``absgrad``
    (``meta["means2d"].absgrad``, read at ``freegaussian_model.py:377``), which plain autograd cannot give."""

    @staticmethod
    def forward(ctx, x, g, sink, col):
        ctx.g, ctx.sink, ctx.col = g, sink, col
        return x.clone()

    @staticmethod
    def backward(ctx, grad):
        ctx.sink[:, ctx.col].index_add_(0, ctx.g, grad.detach().abs().sum(1))
        return grad, None, None, None


def rasterize_to_pixels(
    means2d: Tensor,  # [C,N,2]
    conics: Tensor,  # [C,N,3]
    colors: Tensor,  # [C,N,D]
    opacities: Tensor,  # [C,N]
    width: int,
    height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C,tile_h,tile_w] int32
    flatten_ids: Tensor,  # [M] int32
    backgrounds: Optional[Tensor] = None,  # [C,D]
    flow_affine: Optional[Tensor] = None,  # [C,N,4] row-major 2x2 (covariance flow mode, A.7)
    flow_channels: Optional[Tuple[int, int]] = None,
    chunk: int = 512,
    absgrad_sink: Optional[Tensor] = None,  # [C*N,2], filled during backward with sum_p |d loss / d means2d|
) -> Tuple[Tensor, Tensor, Tensor]:
    """Per-tile front-to-back alpha compositing (Appendix A.6), differentiable.

    Returns render [C,H,W,D], alphas [C,H,W,1], last_ids [C,H,W] int32 (index into the
    sorted list of the last composited Gaussian; 0 when none).

    With ``flow_affine`` the two channels ``flow_channels`` of Gaussian g at pixel p are
    ``colors[g, ch] + A_g (p - mu_g)`` (Appendix A.7 covariance mode).
    """
    C, N, D = colors.shape
    dtype = colors.dtype
    tile_h, tile_w = isect_offsets.shape[1:]
    M = flatten_ids.shape[0]
    offs = isect_offsets.reshape(-1).tolist() + [M]
    m2 = means2d.reshape(C * N, 2)
    cn = conics.reshape(C * N, 3)
    cl = colors.reshape(C * N, D)
    op = opacities.reshape(C * N)
    fa = flow_affine.reshape(C * N, 4) if flow_affine is not None else None
    fid = flatten_ids.to(torch.int64)

    render = torch.zeros(C, height, width, D, dtype=dtype)
    alphas = torch.zeros(C, height, width, 1, dtype=dtype)
    last_ids = torch.zeros(C, height, width, dtype=torch.int32)
    rows, rows_a = [], []  # (index tuple, value) pieces assembled without in-place autograd writes
    for c in range(C):
        for ti in range(tile_h):
            for tj in range(tile_w):
                t = (c * tile_h + ti) * tile_w + tj
                s, e = offs[t], offs[t + 1]
                y0, x0 = ti * tile_size, tj * tile_size
                y1, x1 = min(y0 + tile_size, height), min(x0 + tile_size, width)
                ys = torch.arange(y0, y1, dtype=dtype) + 0.5
                xs = torch.arange(x0, x1, dtype=dtype) + 0.5
                py, px = torch.meshgrid(ys, xs, indexing="ij")
                px, py = px.reshape(-1), py.reshape(-1)
                P = px.shape[0]
                T = torch.ones(P, dtype=dtype)
                acc = torch.zeros(P, D, dtype=dtype)
                stopped = torch.zeros(P, dtype=torch.bool)
                last = torch.zeros(P, dtype=torch.int64)
                for b in range(s, e, chunk):
                    g = fid[b : min(b + chunk, e)]
                    G = g.shape[0]
                    mx, my = m2[g, 0][:, None], m2[g, 1][:, None]
                    if absgrad_sink is not None:
                        mx = _AbsGradTap.apply(mx.expand(G, P), g, absgrad_sink, 0)
                        my = _AbsGradTap.apply(my.expand(G, P), g, absgrad_sink, 1)
                    dx = mx - px[None]
                    dy = my - py[None]
                    con = cn[g]
                    sigma = 0.5 * (con[:, 0:1] * dx * dx + con[:, 2:3] * dy * dy) + con[:, 1:2] * dx * dy
                    alpha = torch.clamp(op[g][:, None] * torch.exp(-sigma), max=ALPHA_MAX)
                    with torch.no_grad():
                        valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
                        a_v = torch.where(valid, alpha, torch.zeros((), dtype=dtype))
                        seq = torch.cat([T.detach()[None], 1 - a_v], 0).cumprod(0)  # seq[g+1] = T after g
                        stop_here = valid & (seq[1:] <= T_STOP)
                        alive = (torch.cumsum(stop_here.to(torch.int64), 0) == 0) & ~stopped[None]
                        use = valid & alive
                    a_m = torch.where(use, alpha, torch.zeros((), dtype=dtype))
                    seq = torch.cat([T[None], 1 - a_m], 0).cumprod(0)
                    w = a_m * seq[:-1]  # alpha * T_before
                    if fa is None:
                        acc = acc + torch.einsum("gp,gd->pd", w, cl[g])
                    else:
                        f0, f1 = flow_channels
                        A = fa[g]
                        extra0 = -(A[:, 0:1] * dx + A[:, 1:2] * dy)  # A (p - mu) = -A delta
                        extra1 = -(A[:, 2:3] * dx + A[:, 3:4] * dy)
                        add = torch.einsum("gp,gd->pd", w, cl[g])
                        e0 = (w * extra0).sum(0)
                        e1 = (w * extra1).sum(0)
                        onehot0 = torch.zeros(D, dtype=dtype)
                        onehot0[f0] = 1
                        onehot1 = torch.zeros(D, dtype=dtype)
                        onehot1[f1] = 1
                        acc = acc + add + e0[:, None] * onehot0 + e1[:, None] * onehot1
                    T = seq[-1]
                    with torch.no_grad():
                        idx = torch.arange(b, b + G)[:, None].expand(G, P)
                        cand = torch.where(use, idx, torch.full_like(idx, -1)).max(0).values
                        last = torch.where(cand >= 0, cand, last)
                        stopped = stopped | stop_here.any(0)
                a_out = 1 - T
                if backgrounds is not None:
                    acc = acc + T[:, None] * backgrounds[c][None]
                rows.append((c, y0, y1, x0, x1, acc.reshape(y1 - y0, x1 - x0, D)))
                rows_a.append(a_out.reshape(y1 - y0, x1 - x0, 1))
                last_ids[c, y0:y1, x0:x1] = last.reshape(y1 - y0, x1 - x0).to(torch.int32)
    # assemble (differentiably) by tile rows
    k = 0
    cams = []
    cams_a = []
    for c in range(C):
        strips, strips_a = [], []
        for ti in range(tile_h):
            strips.append(torch.cat([rows[k + j][5] for j in range(tile_w)], 1))
            strips_a.append(torch.cat([rows_a[k + j] for j in range(tile_w)], 1))
            k += tile_w
        cams.append(torch.cat(strips, 0))
        cams_a.append(torch.cat(strips_a, 0))
    render = torch.stack(cams, 0)
    alphas = torch.stack(cams_a, 0)
    return render, alphas, last_ids


# --------------------------------------------------------------------------- A.7
def cholesky2(cov: Tensor) -> Tensor:
    """Lower Cholesky factor of [[a,b],[b,c]] given as (a,b,c) -> (l00, l10, l11)."""
    a, b, c = cov.unbind(-1)
    l00 = torch.sqrt(a)
    l10 = b / l00
    l11 = torch.sqrt(c - l10 * l10)
    return torch.stack([l00, l10, l11], -1)


def flow_affine_from_cov(cov_t: Tensor, cov_n: Tensor) -> Tensor:
    """A = B(t+1) B(t)^-1 - I for lower-Cholesky B, row-major [.,4] (Appendix A.7)."""
    l00, l10, l11 = cholesky2(cov_t).unbind(-1)
    m00, m10, m11 = cholesky2(cov_n).unbind(-1)
    # B^-1 = [[1/l00, 0], [-l10/(l00 l11), 1/l11]]
    i00 = 1.0 / l00
    i11 = 1.0 / l11
    i10 = -l10 * i00 * i11
    a00 = m00 * i00 - 1.0
    a01 = torch.zeros_like(a00)
    a10 = m10 * i00 + m11 * i10
    a11 = m11 * i11 - 1.0
    return torch.stack([a00, a01, a10, a11], -1)


# --------------------------------------------------------------------------- boundary
def rasterization(
    means: Tensor,
    quats: Tensor,
    scales: Tensor,
    opacities: Tensor,
    colors: Tensor,
    viewmats: Tensor,
    Ks: Tensor,
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = False,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    # --- north_star extension ---
    means_next: Optional[Tensor] = None,
    quats_next: Optional[Tensor] = None,
    scales_next: Optional[Tensor] = None,
    flow_mode: str = "mean",
) -> Tuple[Tensor, Tensor, Dict]:
    """Oracle for the call at ``freegaussian_model.py:847-868`` (Appendix A.1-A.8).

    ``packed`` only changes the layout of ``meta`` (compacted to visible (c,n) pairs in
    ascending order, Appendix A.8); images are identical.
    """
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    assert rasterize_mode in ("classic", "antialiased")
    assert flow_mode in ("mean", "cov")
    C, N = viewmats.shape[0], means.shape[0]
    radii, means2d, depths, conics, comp, cov2d = fully_fused_projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip
    )
    vis = radii > 0
    opac = opacities[None].expand(C, N)
    if rasterize_mode == "antialiased":
        opac = opac * comp

    if sh_degree is None:
        cols = colors[None].expand(C, -1, -1) if colors.dim() == 2 else colors
    else:
        campos = torch.inverse(viewmats)[:, :3, 3]
        dirs = means[None] - campos[:, None]
        shs = colors[None].expand(C, -1, -1, -1) if colors.dim() == 3 else colors
        cols = spherical_harmonics(sh_degree, dirs, shs, masks=vis)
        cols = torch.clamp_min(cols + 0.5, 0.0)

    if render_mode in ("D", "ED"):
        cols = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1, dtype=cols.dtype)
    elif render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[..., None]], -1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, dtype=cols.dtype)], -1)
    n_user = cols.shape[-1]

    flow_affine = None
    flow_channels = None
    if means_next is not None:
        uv_next, z_next = project_points(means_next, viewmats, Ks)
        ok = (vis & (z_next >= near_plane))[..., None]
        flow2d = torch.where(ok, uv_next - means2d, torch.zeros((), dtype=cols.dtype))
        cols = torch.cat([cols, flow2d], -1)
        flow_channels = (n_user, n_user + 1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 2, dtype=cols.dtype)], -1)
        if flow_mode == "cov":
            qn = quats if quats_next is None else quats_next
            sn = scales if scales_next is None else scales_next
            c2n, _, _ = project_cov2d(means_next, qn, sn, viewmats, Ks, width, height)
            eps = torch.tensor([eps2d, 0.0, eps2d], dtype=cols.dtype)
            cov_n = torch.stack([c2n[..., 0, 0], c2n[..., 0, 1], c2n[..., 1, 1]], -1) + eps
            # outside the frame-t visible set, or behind the near plane at t+1: no term
            ident = torch.tensor([1.0, 0.0, 1.0], dtype=cols.dtype)
            safe_t = torch.where(vis[..., None], cov2d, ident)
            safe_n = torch.where(ok, cov_n, ident)
            flow_affine = flow_affine_from_cov(safe_t, safe_n)
            flow_affine = torch.where(ok, flow_affine, torch.zeros((), dtype=cols.dtype))

    tile_w = math.ceil(width / tile_size)
    tile_h = math.ceil(height / tile_size)
    tpg, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tile_w, tile_h)
    isect_offsets = isect_offset_encode(isect_ids, C, tile_w, tile_h)
    absgrad_sink = torch.zeros(C * N, 2, dtype=means2d.dtype) if absgrad else None
    render, alphas, last_ids = rasterize_to_pixels(
        means2d, conics, cols, opac, width, height, tile_size, isect_offsets, flatten_ids,
        backgrounds=backgrounds, flow_affine=flow_affine, flow_channels=flow_channels, absgrad_sink=absgrad_sink,
    )
    flow = None
    if means_next is not None:
        flow = render[..., n_user:]
        render = render[..., :n_user]
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1], render[..., -1:] / alphas.clamp(min=1e-10)], -1)

    meta = {
        "radii": radii, "means2d": means2d, "depths": depths, "conics": conics,
        "opacities": opac, "tile_width": tile_w, "tile_height": tile_h, "tiles_per_gauss": tpg,
        "isect_ids": isect_ids, "flatten_ids": flatten_ids, "isect_offsets": isect_offsets,
        "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
        "last_ids": last_ids, "colors": cols,
    }
    if flow is not None:
        meta["flow"] = flow
    if absgrad_sink is not None:  # complete after backward(): gsplat's means2d.absgrad, [C,N,2]
        meta["absgrad"] = absgrad_sink.view(C, N, 2)
    if packed:
        idx = torch.nonzero(vis.reshape(-1)).squeeze(-1)
        meta["camera_ids"] = (idx // N).to(torch.int64)
        meta["gaussian_ids"] = (idx % N).to(torch.int64)
        for k in ("radii", "means2d", "depths", "conics", "opacities"):
            v = meta[k]
            meta[k] = v.reshape((C * N,) + v.shape[2:])[idx]
    return render, alphas, meta
