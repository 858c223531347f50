// CPU harness for tests/test_host_math.py: runs the per-Gaussian math of
// freegaussian_b200/csrc/splat_math.h (the functions the sm_100a projection kernels call)
// on the host so the hand-derived VJPs can be checked against torch.autograd through the
// oracle without a GPU.  Test infrastructure only -- never linked into the product library.
#include <cstring>

#include "../../freegaussian_b200/csrc/splat_math.h"

using namespace fg;

extern "C" {

void h_project_fwd(int C, int N, const float* means, const float* quats, const float* scales, const float* viewmats,
                   const float* Ks, int W, int H, float eps2d, float near_plane, float far_plane, float radius_clip,
                   int tile_size, int sh_degree, int sh_bases, const float* sh, const float* means_next, int* radii,
                   float* means2d, float* depths, float* conics, float* comps, float* rgb, float* flow, int* tiles) {
    ProjConsts pc = {W, H, eps2d, near_plane, far_plane, radius_clip};
    int tile_w = (W + tile_size - 1) / tile_size, tile_h = (H + tile_size - 1) / tile_size;
    for (int c = 0; c < C; ++c) {
        Camera cam = load_camera(viewmats + 16 * c, Ks + 9 * c);
        for (int n = 0; n < N; ++n) {
            size_t i = (size_t)c * N + n;
            Sym3 cov = quat_scale_to_cov(quats + 4 * n, scales + 3 * n);
            Projected o;
            bool ok = project_gaussian(means + 3 * n, cov, cam, pc, o);
            if (!ok) { o.mx = o.my = o.depth = o.ca = o.cb = o.cc = o.comp = 0.f; }
            radii[i] = o.radius;
            means2d[2 * i] = o.mx; means2d[2 * i + 1] = o.my;
            depths[i] = o.depth;
            conics[3 * i] = o.ca; conics[3 * i + 1] = o.cb; conics[3 * i + 2] = o.cc;
            comps[i] = o.comp;
            int nt = 0;
            float fu = 0.f, fv = 0.f;
            float col[3] = {0.f, 0.f, 0.f};
            if (ok) {
                TileRect r = tile_rect(o.mx, o.my, o.radius, tile_size, tile_w, tile_h);
                nt = (r.x1 - r.x0) * (r.y1 - r.y0);
                if (means_next) {
                    float u, v;
                    if (project_point(means_next + 3 * n, cam, near_plane, u, v)) { fu = u - o.mx; fv = v - o.my; }
                }
                if (sh_degree >= 0) {
                    const float* m = means + 3 * n;
                    float dx = m[0] - cam.pos[0], dy = m[1] - cam.pos[1], dz = m[2] - cam.pos[2];
                    float inorm = rsqrt_f(dx * dx + dy * dy + dz * dz);
                    float B[16];
                    sh_basis(sh_degree, dx * inorm, dy * inorm, dz * inorm, B);
                    const float* co = sh + (size_t)n * sh_bases * 3;
                    for (int k = 0; k < (sh_degree + 1) * (sh_degree + 1); ++k)
                        for (int ch = 0; ch < 3; ++ch) col[ch] += B[k] * co[3 * k + ch];
                    for (int ch = 0; ch < 3; ++ch) col[ch] = fmaxf_(col[ch] + 0.5f, 0.f);
                }
            }
            tiles[i] = nt;
            if (rgb) { rgb[3 * i] = col[0]; rgb[3 * i + 1] = col[1]; rgb[3 * i + 2] = col[2]; }
            if (flow) { flow[2 * i] = fu; flow[2 * i + 1] = fv; }
        }
    }
}

void h_project_bwd(int C, int N, const float* means, const float* quats, const float* scales, const float* viewmats,
                   const float* Ks, int W, int H, float eps2d, float near_plane, float far_plane, float radius_clip,
                   int sh_degree, int sh_bases, const float* sh, const float* means_next, const int* radii,
                   const float* v_means2d, const float* v_depths, const float* v_conics, const float* v_comps,
                   const float* v_rgb, const float* v_flow, float* v_means, float* v_quats, float* v_scales,
                   float* v_sh, float* v_means_next) {
    ProjConsts pc = {W, H, eps2d, near_plane, far_plane, radius_clip};
    memset(v_means, 0, sizeof(float) * 3 * N);
    memset(v_quats, 0, sizeof(float) * 4 * N);
    memset(v_scales, 0, sizeof(float) * 3 * N);
    if (v_sh) memset(v_sh, 0, sizeof(float) * 3 * sh_bases * N);
    if (v_means_next) memset(v_means_next, 0, sizeof(float) * 3 * N);
    for (int n = 0; n < N; ++n) {
        const float* m = means + 3 * n;
        Sym3 cov = quat_scale_to_cov(quats + 4 * n, scales + 3 * n);
        Sym3 G = {};
        for (int c = 0; c < C; ++c) {
            size_t i = (size_t)c * N + n;
            if (radii[i] <= 0) continue;
            Camera cam = load_camera(viewmats + 16 * c, Ks + 9 * c);
            float v_m2d[2] = {v_means2d[2 * i], v_means2d[2 * i + 1]};
            float v_con[3] = {v_conics[3 * i], v_conics[3 * i + 1], v_conics[3 * i + 2]};
            float v_depth = v_depths[i], v_comp = v_comps ? v_comps[i] : 0.f;
            if (v_flow && means_next) {
                float u, v;
                if (project_point(means_next + 3 * n, cam, near_plane, u, v)) {
                    v_m2d[0] -= v_flow[2 * i]; v_m2d[1] -= v_flow[2 * i + 1];
                    project_point_vjp(means_next + 3 * n, cam, v_flow[2 * i], v_flow[2 * i + 1], v_means_next + 3 * n);
                }
            }
            if (sh_degree >= 0 && v_rgb) {
                float dx = m[0] - cam.pos[0], dy = m[1] - cam.pos[1], dz = m[2] - cam.pos[2];
                float inorm = rsqrt_f(dx * dx + dy * dy + dz * dz);
                float x = dx * inorm, y = dy * inorm, z = dz * inorm;
                float B[16];
                sh_basis(sh_degree, x, y, z, B);
                const float* co = sh + (size_t)n * sh_bases * 3;
                float col[3] = {0.f, 0.f, 0.f};
                int K = (sh_degree + 1) * (sh_degree + 1);
                for (int k = 0; k < K; ++k)
                    for (int ch = 0; ch < 3; ++ch) col[ch] += B[k] * co[3 * k + ch];
                float vr[3];
                for (int ch = 0; ch < 3; ++ch) vr[ch] = (col[ch] + 0.5f > 0.f) ? v_rgb[3 * i + ch] : 0.f;
                float sk[16];
                for (int k = 0; k < K; ++k) {
                    for (int ch = 0; ch < 3; ++ch) v_sh[((size_t)n * sh_bases + k) * 3 + ch] += B[k] * vr[ch];
                    sk[k] = co[3 * k] * vr[0] + co[3 * k + 1] * vr[1] + co[3 * k + 2] * vr[2];
                }
                float vd[3];
                sh_basis_vjp(sh_degree, x, y, z, sk, vd);
                float dot = vd[0] * x + vd[1] * y + vd[2] * z;
                v_means[3 * n] += (vd[0] - dot * x) * inorm;
                v_means[3 * n + 1] += (vd[1] - dot * y) * inorm;
                v_means[3 * n + 2] += (vd[2] - dot * z) * inorm;
            }
            project_gaussian_vjp(m, cov, cam, pc, v_m2d, v_depth, v_con, v_comp, v_means + 3 * n, G);
        }
        quat_scale_to_cov_vjp(quats + 4 * n, scales + 3 * n, G, v_quats + 4 * n, v_scales + 3 * n);
    }
}

void h_flow_affine(int n, const float* ct, const float* cn, float* A) {
    for (int i = 0; i < n; ++i) flow_affine(ct + 3 * i, cn + 3 * i, A + 4 * i);
}
void h_flow_affine_vjp(int n, const float* ct, const float* cn, const float* vA, float* v_ct, float* v_cn) {
    for (int i = 0; i < n; ++i) {
        v_ct[3 * i] = v_ct[3 * i + 1] = v_ct[3 * i + 2] = 0.f;
        v_cn[3 * i] = v_cn[3 * i + 1] = v_cn[3 * i + 2] = 0.f;
        flow_affine_vjp(ct + 3 * i, cn + 3 * i, vA + 4 * i, v_ct + 3 * i, v_cn + 3 * i);
    }
}

}  // extern "C"
