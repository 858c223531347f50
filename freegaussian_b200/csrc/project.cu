// (1) Fused 3D->2D projection: quat/scale -> covariance, world->camera, EWA Jacobian,
// blur, conic, radius, near/far + frustum culling, SH->RGB (+0.5, clamp), depth and flow
// feature channels, and the per-splat tile count -- one pass over the Gaussian records.
// Forward and backward.  Replaces gsplat fully_fused_projection{,_bwd} +
// spherical_harmonics{,_bwd} behind freegaussian_model.py:847-868.
//
// Roofline: HBM.  Algorithmic bytes per (camera, Gaussian) at SH degree 3:
//   fwd 276 B = 44 (mean, quat, scale) + 192 (SH) + 40 (radii, mean2d, depth, conic, rgb)
//   bwd 548 B = 236 + 40 + 36 read, 236 written                       
//
// Layout: one thread per Gaussian, inner loop over the cameras of this rank (C is 1..4 per
// GPU under view sharding), so the 236-byte parameter record is read once and gradients
// of all cameras are summed in registers -- no atomics, deterministic.  The 192-byte SH
// rows are staged through shared memory with fully coalesced 16-byte loads/stores
// (row stride padded to an odd number of float4 so per-thread row reads are
// bank-conflict free).
#include "common.cuh"
#include "splat_math.h"

namespace fg {

constexpr int PB = 128;  // threads (= Gaussians) per block

struct ProjParams {
    int C, N;
    const float* means;
    const float* quats;
    const float* scales;
    const float* viewmats;
    const float* Ks;
    ProjConsts pc;
    int tile_size, tile_w, tile_h;
    int sh_row_floats;  // floats per Gaussian row of sh_coeffs (sh_bases*3)
    const float* sh;
    const float* means_next;
    const float* quats_next;   // covariance flow mode: frame t+1 rotation / scale (NULL = frame t's)
    const float* scales_next;
    int flow_cov;
    int feat_stride, rgb_off, depth_off, flow_off;
    // forward outputs
    int32_t* radii;
    float* means2d;
    float* depths;
    float* conics;
    float* comps;
    float* feat;
    float* flow_affine;
    int32_t* tiles_per_gauss;
    // backward inputs
    const int32_t* radii_in;
    const float* v_means2d;
    const float* v_depths;
    const float* v_conics;
    const float* v_comps;
    const float* v_feat;
    const float* v_flow_affine;
    const float* feat_fwd;      // forward `feat` (its rgb channels give the clamp mask of the SH colours)
    const float* v_mean_extra;  // [N,3] added to v_means (the SH direction term when SH runs in its own kernel)
    // backward outputs
    float* v_means;
    float* v_quats;
    float* v_scales;
    float* v_sh;
    float* v_means_next;
    float* v_quats_next;
    float* v_scales_next;
    // multi-GPU exchange (csrc/exchange.cu): the SH kernel publishes the clamp-masked colour gradient of every visible
    // (view, Gaussian), a visibility bit mask and the camera centres instead of writing v_sh rows
    float* pub_campos;    // [C,4], followed by one uint32: nnz
    uint32_t* pub_mask;   // [C,pub_words]
    uint32_t* pub_prefix; // [C,pub_words]: compact row of the first visible splat of each mask word
    float* pub_rgb;       // [nnz,3] compact, ascending c*N+n
    const int32_t* pub_offs;     // [C*N] exclusive scan of (radii > 0)
    const long long* pub_nnz;    // its total (device)
    int pub_words;
};

template <int DEG>
struct ShShape {
    static constexpr int NEED = 3 * (DEG + 1) * (DEG + 1);  // floats used per row
    static constexpr int NV = (NEED + 3) / 4;               // float4 per row staged
    static constexpr int ROWV = NV | 1;                     // odd row stride (float4)
    static constexpr int ROWF = NEED | 1;                   // odd row stride (float), unaligned path
};

template <int DEG, bool VEC4>
__device__ __forceinline__ void stage_sh_rows(const float* __restrict__ sh, int row_floats, int n0, int N,
                                              float* smem) {
    using S = ShShape<DEG>;
    const int rows = min(PB, N - n0);
    if (VEC4) {
        const float4* src = reinterpret_cast<const float4*>(sh);
        const int gv = row_floats >> 2;
        float4* dst = reinterpret_cast<float4*>(smem);
        for (int q = threadIdx.x; q < rows * S::NV; q += PB) {
            int g = q / S::NV, j = q - g * S::NV;
            dst[g * S::ROWV + j] = __ldg(src + (size_t)(n0 + g) * gv + j);
        }
    } else {
        for (int q = threadIdx.x; q < rows * S::NEED; q += PB) {
            int g = q / S::NEED, k = q - g * S::NEED;
            smem[g * S::ROWF + k] = __ldg(sh + (size_t)(n0 + g) * row_floats + k);
        }
    }
}

template <int DEG, bool VEC4>
__device__ __forceinline__ void read_sh_row(const float* smem, float* coef) {
    using S = ShShape<DEG>;
    if (VEC4) {
        const float4* r = reinterpret_cast<const float4*>(smem) + threadIdx.x * S::ROWV;
#pragma unroll
        for (int j = 0; j < S::NV; ++j) {
            float4 v = r[j];
            coef[4 * j] = v.x; coef[4 * j + 1] = v.y; coef[4 * j + 2] = v.z; coef[4 * j + 3] = v.w;
        }
    } else {
        const float* r = smem + threadIdx.x * S::ROWF;
#pragma unroll
        for (int k = 0; k < S::NEED; ++k) coef[k] = r[k];
    }
}

// Sparse visibility: a thread fetches just its own row straight from global memory (no staging),
// so rows of culled Gaussians are never read.
template <int DEG, bool VEC4>
__device__ __forceinline__ void read_sh_row_global(const float* __restrict__ sh, int row_floats, int n, float* coef) {
    using S = ShShape<DEG>;
    if (VEC4) {
        const float4* r = reinterpret_cast<const float4*>(sh) + (size_t)n * (row_floats >> 2);
#pragma unroll
        for (int j = 0; j < S::NV; ++j) {
            float4 v = __ldg(r + j);
            coef[4 * j] = v.x; coef[4 * j + 1] = v.y; coef[4 * j + 2] = v.z; coef[4 * j + 3] = v.w;
        }
    } else {
        const float* r = sh + (size_t)n * row_floats;
#pragma unroll
        for (int k = 0; k < S::NEED; ++k) coef[k] = __ldg(r + k);
    }
}

template <int DEG, bool VEC4>
__global__ void __launch_bounds__(PB, 4) project_fwd_kernel(ProjParams p) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    const int n0 = blockIdx.x * PB;
    const int n = n0 + threadIdx.x;
    const bool in_range = n < p.N;

    float m[3] = {0.f, 0.f, 0.f}, mn[3] = {0.f, 0.f, 0.f};
    Sym3 cov = {}, cov_next = {};
    if (in_range) {
        float q[4], s[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i] = __ldg(p.means + 3 * (size_t)n + i);
        float4 qq = __ldg(reinterpret_cast<const float4*>(p.quats) + n);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
#pragma unroll
        for (int i = 0; i < 3; ++i) s[i] = __ldg(p.scales + 3 * (size_t)n + i);
        cov = quat_scale_to_cov(q, s);
        if (p.means_next) {
#pragma unroll
            for (int i = 0; i < 3; ++i) mn[i] = __ldg(p.means_next + 3 * (size_t)n + i);
        }
        if (p.flow_cov) {
            float qn[4] = {q[0], q[1], q[2], q[3]}, sn[3] = {s[0], s[1], s[2]};
            if (p.quats_next) {
                float4 t = __ldg(reinterpret_cast<const float4*>(p.quats_next) + n);
                qn[0] = t.x; qn[1] = t.y; qn[2] = t.z; qn[3] = t.w;
            }
            if (p.scales_next) {
#pragma unroll
                for (int i = 0; i < 3; ++i) sn[i] = __ldg(p.scales_next + 3 * (size_t)n + i);
            }
            cov_next = quat_scale_to_cov(qn, sn);
        }
    }

    // SH rows: staged through shared memory (coalesced) when at least half of the block is visible,
    // otherwise each visible thread reads its own row and culled rows are never touched
    int sh_mode = 0;  // 0 undecided, 1 staged, 2 per-thread
    bool have_coef = false;
    float coef[ShShape<(DEG >= 0 ? DEG : 0)>::NEED];
    for (int c0 = 0; c0 < p.C; c0 += 32) {
        const int c1 = min(p.C, c0 + 32);
        uint32_t vis = 0;
        for (int c = c0; c < c1; ++c) {
            if (!in_range) continue;
            const Camera cam = load_camera(p.viewmats + 16 * c, p.Ks + 9 * c);
            Projected o;
            const bool ok = project_gaussian(m, cov, cam, p.pc, o);
            const size_t i = (size_t)c * p.N + n;
            int ntiles = 0;
            float fu = 0.f, fv = 0.f;
            float A[4] = {0.f, 0.f, 0.f, 0.f};
            if (ok) {
                vis |= 1u << (c - c0);
                TileRect r = tile_rect(o.mx, o.my, o.radius, p.tile_size, p.tile_w, p.tile_h);
                ntiles = (r.x1 - r.x0) * (r.y1 - r.y0);
                if (p.means_next) {
                    float u, v;
                    if (p.flow_cov) {
                        float cn[3];
                        if (project_cov2d(mn, cov_next, cam, p.pc, cn[0], cn[1], cn[2], u, v)) {
                            fu = u - o.mx; fv = v - o.my;
                            const float ct[3] = {o.a, o.b, o.c};
                            flow_affine(ct, cn, A);
                        }
                    } else if (project_point(mn, cam, p.pc.near_plane, u, v)) { fu = u - o.mx; fv = v - o.my; }
                }
            } else {
                o.mx = o.my = o.depth = o.ca = o.cb = o.cc = o.comp = 0.f;
            }
            p.radii[i] = o.radius;
            reinterpret_cast<float2*>(p.means2d)[i] = make_float2(o.mx, o.my);
            p.depths[i] = o.depth;
            p.conics[3 * i] = o.ca; p.conics[3 * i + 1] = o.cb; p.conics[3 * i + 2] = o.cc;
            if (p.comps) p.comps[i] = o.comp;
            p.tiles_per_gauss[i] = ntiles;
            float* f = p.feat + i * p.feat_stride;
            if (p.depth_off >= 0) f[p.depth_off] = o.depth;
            if (p.flow_off >= 0) { f[p.flow_off] = fu; f[p.flow_off + 1] = fv; }
            if (p.flow_affine) reinterpret_cast<float4*>(p.flow_affine)[i] = make_float4(A[0], A[1], A[2], A[3]);
        }
        if (DEG >= 0) {
            using S = ShShape<(DEG >= 0 ? DEG : 0)>;
            if (sh_mode == 0) {
                // uniform branch: sh_mode only changes under a block-wide vote
                const int nvis = __syncthreads_count(vis != 0);
                if (nvis * 2 >= PB) {
                    stage_sh_rows<(DEG >= 0 ? DEG : 0), VEC4>(p.sh, p.sh_row_floats, n0, p.N, smem);
                    __syncthreads();
                    sh_mode = 1;
                } else if (nvis > 0) {
                    sh_mode = 2;
                }
            }
            if (vis && !have_coef) {
                if (sh_mode == 1) read_sh_row<(DEG >= 0 ? DEG : 0), VEC4>(smem, coef);
                else read_sh_row_global<(DEG >= 0 ? DEG : 0), VEC4>(p.sh, p.sh_row_floats, n, coef);
                have_coef = true;
            }
            for (int c = c0; c < c1; ++c) {
                if (!in_range) continue;
                float rgb[3] = {0.f, 0.f, 0.f};
                if (vis & (1u << (c - c0))) {
                    const Camera cam = load_camera(p.viewmats + 16 * c, p.Ks + 9 * c);
                    float dx = m[0] - cam.pos[0], dy = m[1] - cam.pos[1], dz = m[2] - cam.pos[2];
                    float inorm = rsqrt_f(dx * dx + dy * dy + dz * dz);
                    float B[16];
                    sh_basis(DEG, dx * inorm, dy * inorm, dz * inorm, B);
#pragma unroll
                    for (int k = 0; k < (DEG + 1) * (DEG + 1); ++k) {
                        rgb[0] += B[k] * coef[3 * k];
                        rgb[1] += B[k] * coef[3 * k + 1];
                        rgb[2] += B[k] * coef[3 * k + 2];
                    }
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) rgb[ch] = fmaxf(rgb[ch] + 0.5f, 0.f);
                }
                float* f = p.feat + ((size_t)c * p.N + n) * p.feat_stride + p.rgb_off;
                f[0] = rgb[0]; f[1] = rgb[1]; f[2] = rgb[2];
            }
        }
    }
}

// ------------------------------------------------------------------------------ backward
template <int DEG, bool VEC4, int MINB = (DEG < 0 ? 4 : 3)>
__global__ void __launch_bounds__(PB, MINB) project_bwd_kernel(ProjParams p) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    using S = ShShape<(DEG >= 0 ? DEG : 0)>;
    const int n0 = blockIdx.x * PB;
    const int n = n0 + threadIdx.x;
    const bool in_range = n < p.N;
    constexpr int NEED = (DEG >= 0) ? S::NEED : 1;

    float m[3] = {0.f, 0.f, 0.f}, mn[3] = {0.f, 0.f, 0.f}, q[4] = {1.f, 0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
    float qn[4] = {1.f, 0.f, 0.f, 0.f}, sn[3] = {1.f, 1.f, 1.f};
    Sym3 cov = {}, cov_next = {};
    if (in_range) {
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i] = __ldg(p.means + 3 * (size_t)n + i);
        float4 qq = __ldg(reinterpret_cast<const float4*>(p.quats) + n);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
#pragma unroll
        for (int i = 0; i < 3; ++i) s[i] = __ldg(p.scales + 3 * (size_t)n + i);
        cov = quat_scale_to_cov(q, s);
        if (p.means_next) {
#pragma unroll
            for (int i = 0; i < 3; ++i) mn[i] = __ldg(p.means_next + 3 * (size_t)n + i);
        }
        if (p.flow_cov) {
#pragma unroll
            for (int i = 0; i < 4; ++i) qn[i] = q[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) sn[i] = s[i];
            if (p.quats_next) {
                float4 t = __ldg(reinterpret_cast<const float4*>(p.quats_next) + n);
                qn[0] = t.x; qn[1] = t.y; qn[2] = t.z; qn[3] = t.w;
            }
            if (p.scales_next) {
#pragma unroll
                for (int i = 0; i < 3; ++i) sn[i] = __ldg(p.scales_next + 3 * (size_t)n + i);
            }
            cov_next = quat_scale_to_cov(qn, sn);
        }
    }
    float coef[NEED];
    float vcoef[NEED];
#pragma unroll
    for (int k = 0; k < NEED; ++k) vcoef[k] = 0.f;
    if (DEG >= 0) {
        // same hybrid as the forward: staged when the block is mostly visible, per-thread otherwise
        bool vis_any = false;
        if (in_range)
            for (int c = 0; c < p.C; ++c) vis_any |= p.radii_in[(size_t)c * p.N + n] > 0;
        const int nvis = __syncthreads_count(vis_any);
        if (nvis * 2 >= PB) {
            stage_sh_rows<(DEG >= 0 ? DEG : 0), VEC4>(p.sh, p.sh_row_floats, n0, p.N, smem);
            __syncthreads();
            if (vis_any) read_sh_row<(DEG >= 0 ? DEG : 0), VEC4>(smem, coef);
        } else if (vis_any) {
            read_sh_row_global<(DEG >= 0 ? DEG : 0), VEC4>(p.sh, p.sh_row_floats, n, coef);
        }
    }

    float v_mean[3] = {0.f, 0.f, 0.f}, v_mean_next[3] = {0.f, 0.f, 0.f};
    Sym3 G = {}, G_next = {};
    for (int c = 0; c < p.C; ++c) {
        if (!in_range) break;
        const size_t i = (size_t)c * p.N + n;
        if (p.radii_in[i] <= 0) continue;
        const Camera cam = load_camera(p.viewmats + 16 * c, p.Ks + 9 * c);
        float v_m2d[2] = {0.f, 0.f}, v_con[3] = {0.f, 0.f, 0.f};
        float v_depth = 0.f, v_comp = 0.f;
        if (p.v_means2d) { float2 t = reinterpret_cast<const float2*>(p.v_means2d)[i]; v_m2d[0] = t.x; v_m2d[1] = t.y; }
        if (p.v_conics) { v_con[0] = p.v_conics[3 * i]; v_con[1] = p.v_conics[3 * i + 1]; v_con[2] = p.v_conics[3 * i + 2]; }
        if (p.v_depths) v_depth = p.v_depths[i];
        if (p.v_comps) v_comp = p.v_comps[i];
        if (p.v_feat) {
            const float* vf = p.v_feat + i * p.feat_stride;
            if (p.depth_off >= 0) v_depth += vf[p.depth_off];
            if (p.flow_off >= 0 && p.means_next && !p.flow_cov) {
                float u, v;
                if (project_point(mn, cam, p.pc.near_plane, u, v)) {
                    float vu = vf[p.flow_off], vv = vf[p.flow_off + 1];
                    v_m2d[0] -= vu; v_m2d[1] -= vv;
                    project_point_vjp(mn, cam, vu, vv, v_mean_next);
                }
            }
            if (DEG >= 0) {
                float dx = m[0] - cam.pos[0], dy = m[1] - cam.pos[1], dz = m[2] - cam.pos[2];
                float inorm = rsqrt_f(dx * dx + dy * dy + dz * dz);
                float x = dx * inorm, y = dy * inorm, z = dz * inorm;
                float B[16];
                sh_basis(DEG, x, y, z, B);
                float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < (DEG + 1) * (DEG + 1); ++k) {
                    rgb[0] += B[k] * coef[3 * k]; rgb[1] += B[k] * coef[3 * k + 1]; rgb[2] += B[k] * coef[3 * k + 2];
                }
                float vr[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) vr[ch] = (rgb[ch] + 0.5f > 0.f) ? vf[p.rgb_off + ch] : 0.f;
                float sk[16];
#pragma unroll
                for (int k = 0; k < (DEG + 1) * (DEG + 1); ++k) {
                    vcoef[3 * k] += B[k] * vr[0]; vcoef[3 * k + 1] += B[k] * vr[1]; vcoef[3 * k + 2] += B[k] * vr[2];
                    sk[k] = coef[3 * k] * vr[0] + coef[3 * k + 1] * vr[1] + coef[3 * k + 2] * vr[2];
                }
                float vd[3];
                sh_basis_vjp(DEG, x, y, z, sk, vd);
                float dot = vd[0] * x + vd[1] * y + vd[2] * z;
                v_mean[0] += (vd[0] - dot * x) * inorm;
                v_mean[1] += (vd[1] - dot * y) * inorm;
                v_mean[2] += (vd[2] - dot * z) * inorm;
            }
        }
        float v_ct[3] = {0.f, 0.f, 0.f};
        bool have_ct = false;
        if (p.flow_cov && p.means_next) {
            // covariance flow mode: the frame t+1 splat contributes through its mean AND its 2-D covariance
            float ct[3], cn[3], ut, vt, un, vn;
            if (project_cov2d(mn, cov_next, cam, p.pc, cn[0], cn[1], cn[2], un, vn) &&
                project_cov2d(m, cov, cam, p.pc, ct[0], ct[1], ct[2], ut, vt)) {
                float vuv[2] = {0.f, 0.f};
                if (p.v_feat && p.flow_off >= 0) {
                    const float* vf = p.v_feat + i * p.feat_stride;
                    vuv[0] = vf[p.flow_off]; vuv[1] = vf[p.flow_off + 1];
                    v_m2d[0] -= vuv[0]; v_m2d[1] -= vuv[1];
                }
                float v_cn[3] = {0.f, 0.f, 0.f};
                if (p.v_flow_affine) {
                    const float4 t = reinterpret_cast<const float4*>(p.v_flow_affine)[i];
                    const float vA[4] = {t.x, t.y, t.z, t.w};
                    flow_affine_vjp(ct, cn, vA, v_ct, v_cn);
                    have_ct = true;
                }
                const float zero3[3] = {0.f, 0.f, 0.f};
                project_gaussian_vjp(mn, cov_next, cam, p.pc, vuv, 0.f, zero3, 0.f, v_mean_next, G_next, v_cn);
            }
        }
        project_gaussian_vjp(m, cov, cam, p.pc, v_m2d, v_depth, v_con, v_comp, v_mean, G, have_ct ? v_ct : nullptr);
    }
    if (in_range) {
        float v_q[4] = {0.f, 0.f, 0.f, 0.f}, v_s[3] = {0.f, 0.f, 0.f};
        quat_scale_to_cov_vjp(q, s, G, v_q, v_s);
        if (p.flow_cov) {
            // split G_next between rotation and scale of frame t+1; tensors that were not given
            // separately are frame t's, so their share folds into v_q / v_s
            float v_qn[4] = {0.f, 0.f, 0.f, 0.f}, v_sn[3] = {0.f, 0.f, 0.f};
            quat_scale_to_cov_vjp(qn, sn, G_next, v_qn, v_sn);
            if (p.v_quats_next) reinterpret_cast<float4*>(p.v_quats_next)[n] = make_float4(v_qn[0], v_qn[1], v_qn[2], v_qn[3]);
            else { v_q[0] += v_qn[0]; v_q[1] += v_qn[1]; v_q[2] += v_qn[2]; v_q[3] += v_qn[3]; }
            if (p.v_scales_next) {
#pragma unroll
                for (int i = 0; i < 3; ++i) p.v_scales_next[3 * (size_t)n + i] = v_sn[i];
            } else { v_s[0] += v_sn[0]; v_s[1] += v_sn[1]; v_s[2] += v_sn[2]; }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
            p.v_means[3 * (size_t)n + i] = v_mean[i] + (p.v_mean_extra ? p.v_mean_extra[3 * (size_t)n + i] : 0.f);
        reinterpret_cast<float4*>(p.v_quats)[n] = make_float4(v_q[0], v_q[1], v_q[2], v_q[3]);
#pragma unroll
        for (int i = 0; i < 3; ++i) p.v_scales[3 * (size_t)n + i] = v_s[i];
        if (p.v_means_next) {
#pragma unroll
            for (int i = 0; i < 3; ++i) p.v_means_next[3 * (size_t)n + i] = v_mean_next[i];
        }
    }
    if (DEG >= 0) {
        // write v_sh rows through shared memory so the global stores are coalesced.
        // Full rows are written (zeros beyond the evaluated bases), so no pre-zeroing is needed.
        __syncthreads();  // everyone has read its coefficient row
        const int row_floats = p.sh_row_floats;
        const int rows = min(PB, p.N - n0);
        // pass over the row in chunks of NEED floats held in smem (first chunk = real grads)
        if (VEC4) {
            float4* r = reinterpret_cast<float4*>(smem) + threadIdx.x * S::ROWV;
#pragma unroll
            for (int j = 0; j < S::NV; ++j) {
                float4 v;
                v.x = (4 * j < NEED) ? vcoef[4 * j] : 0.f;
                v.y = (4 * j + 1 < NEED) ? vcoef[4 * j + 1] : 0.f;
                v.z = (4 * j + 2 < NEED) ? vcoef[4 * j + 2] : 0.f;
                v.w = (4 * j + 3 < NEED) ? vcoef[4 * j + 3] : 0.f;
                r[j] = v;
            }
            __syncthreads();
            const int gv = row_floats >> 2;
            float4* dst = reinterpret_cast<float4*>(p.v_sh);
            const float4* src = reinterpret_cast<const float4*>(smem);
            for (int qd = threadIdx.x; qd < rows * gv; qd += PB) {
                int g = qd / gv, j = qd - g * gv;
                float4 v = (j < S::NV) ? src[g * S::ROWV + j] : make_float4(0.f, 0.f, 0.f, 0.f);
                dst[(size_t)(n0 + g) * gv + j] = v;
            }
        } else {
            float* r = smem + threadIdx.x * S::ROWF;
#pragma unroll
            for (int k = 0; k < NEED; ++k) r[k] = vcoef[k];
            __syncthreads();
            for (int qd = threadIdx.x; qd < rows * row_floats; qd += PB) {
                int g = qd / row_floats, k = qd - g * row_floats;
                p.v_sh[(size_t)(n0 + g) * row_floats + k] = (k < NEED) ? smem[g * S::ROWF + k] : 0.f;
            }
        }
    }
}

// ------------------------------------------------------------------------------ SH backward
// The spherical-harmonics half of the backward pass as its own streaming kernel (16-base rows,
// 16-byte aligned): v_sh rows are accumulated in shared memory and written with coalesced
// float4 stores; coefficients are consumed four bases (three float4) at a time straight from a
// staged copy or from global memory, so no 48-float arrays live in registers (the fused kernel
// needed 168-250 registers and ran at 12-18 % occupancy).  The clamp mask of max(rgb + 0.5, 0)
// comes from the forward colours in `feat`.  Also writes the direction term of v_means, which
// the geometry kernel (project_bwd_kernel<-1>) then adds to its own.
constexpr bool SH_BWD_STAGE = false;  // coefficient rows are read once per camera straight from global memory
template <int DEG>
__global__ void __launch_bounds__(PB) sh_bwd_kernel(ProjParams p) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int NG = (K + 3) / 4;   // groups of four bases
    constexpr int NV3 = NG * 3;       // float4 per row that carry gradient
    constexpr int ROW = 13;           // odd float4 row stride: conflict-free own-row access
    float4* acc = reinterpret_cast<float4*>(smem) + threadIdx.x * ROW;
    float4* stg_all = reinterpret_cast<float4*>(smem) + PB * ROW;
    const int n0 = blockIdx.x * PB;
    const int n = n0 + threadIdx.x;
    const bool in_range = n < p.N;
    const int gv = p.sh_row_floats >> 2;
#pragma unroll
    for (int j = 0; j < NV3; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool vis_any = false;
    const bool want_sh = p.v_sh != nullptr;
    for (int c = 0; c < p.C; ++c) {
        const bool vis = in_range && p.v_feat && p.radii_in[(size_t)c * p.N + n] > 0;
        vis_any |= vis;
        if (p.pub_mask) {  // n0 is a multiple of 128: every warp owns whole mask words
            const uint32_t bits = __ballot_sync(0xffffffffu, vis);
            if ((threadIdx.x & 31) == 0 && n < p.N) {
                p.pub_mask[(size_t)c * p.pub_words + (n >> 5)] = bits;
                p.pub_prefix[(size_t)c * p.pub_words + (n >> 5)] = (uint32_t)p.pub_offs[(size_t)c * p.N + n];
            }
        }
    }
    if (p.pub_campos && blockIdx.x == 0) {
        if (threadIdx.x < p.C) {
            const Camera cam = load_camera(p.viewmats + 16 * threadIdx.x, p.Ks + 9 * threadIdx.x);
            reinterpret_cast<float4*>(p.pub_campos)[threadIdx.x] = make_float4(cam.pos[0], cam.pos[1], cam.pos[2], 0.f);
        }
        if (threadIdx.x == 0) reinterpret_cast<uint32_t*>(p.pub_campos + 4 * p.C)[0] = (uint32_t)*p.pub_nnz;
    }
    const int nvis = __syncthreads_count(vis_any);
    const bool staged = SH_BWD_STAGE && nvis * 2 >= PB;
    if (staged) {
        const float4* src = reinterpret_cast<const float4*>(p.sh);
        const int rows = min(PB, p.N - n0);
        for (int q = threadIdx.x; q < rows * NV3; q += PB) {
            const int g = q / NV3, j = q - g * NV3;
            stg_all[g * ROW + j] = __ldg(src + (size_t)(n0 + g) * gv + j);
        }
        __syncthreads();
    }
    float v_md[3] = {0.f, 0.f, 0.f};
    if (vis_any) {
        float m[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i] = __ldg(p.means + 3 * (size_t)n + i);
        const float4* crow = staged ? stg_all + threadIdx.x * ROW
                                    : reinterpret_cast<const float4*>(p.sh) + (size_t)n * gv;
        for (int c = 0; c < p.C; ++c) {
            const size_t i = (size_t)c * p.N + n;
            if (p.radii_in[i] <= 0) continue;
            const Camera cam = load_camera(p.viewmats + 16 * c, p.Ks + 9 * c);
            const float dx = m[0] - cam.pos[0], dy = m[1] - cam.pos[1], dz = m[2] - cam.pos[2];
            const float inorm = rsqrt_f(dx * dx + dy * dy + dz * dz);
            const float x = dx * inorm, y = dy * inorm, z = dz * inorm;
            float B[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) B[k] = 0.f;
            sh_basis(DEG, x, y, z, B);
            const float* ff = p.feat_fwd + i * p.feat_stride + p.rgb_off;
            const float* vf = p.v_feat + i * p.feat_stride + p.rgb_off;
            float vr[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) vr[ch] = (ff[ch] > 0.f) ? vf[ch] : 0.f;
            if (p.pub_rgb) {
                float* d = p.pub_rgb + (size_t)p.pub_offs[i] * 3;
                d[0] = vr[0]; d[1] = vr[1]; d[2] = vr[2];
            }
            float sk[16];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float f[12];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4 v = staged ? crow[3 * g + j] : __ldg(crow + 3 * g + j);
                    f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) sk[4 * g + b] = f[3 * b] * vr[0] + f[3 * b + 1] * vr[1] + f[3 * b + 2] * vr[2];
                if (want_sh) {
                    float a[12];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float4 v = acc[3 * g + j];
                        a[4 * j] = v.x; a[4 * j + 1] = v.y; a[4 * j + 2] = v.z; a[4 * j + 3] = v.w;
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int k = 4 * g + b;
                        a[3 * b] = fmaf(B[k], vr[0], a[3 * b]);
                        a[3 * b + 1] = fmaf(B[k], vr[1], a[3 * b + 1]);
                        a[3 * b + 2] = fmaf(B[k], vr[2], a[3 * b + 2]);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc[3 * g + j] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                }
            }
            float vd[3];
            sh_basis_vjp(DEG, x, y, z, sk, vd);
            const float dot = vd[0] * x + vd[1] * y + vd[2] * z;
            v_md[0] += (vd[0] - dot * x) * inorm;
            v_md[1] += (vd[1] - dot * y) * inorm;
            v_md[2] += (vd[2] - dot * z) * inorm;
        }
    }
    if (in_range) {
#pragma unroll
        for (int i = 0; i < 3; ++i) p.v_means[3 * (size_t)n + i] = v_md[i];
    }
    if (!want_sh) return;  // exchange mode: the rows are summed over ALL ranks' views by sh_bwd_views_kernel
    __syncthreads();
    // coalesced write of full rows (zeros beyond the evaluated bases)
    const int rows = min(PB, p.N - n0);
    float4* dst = reinterpret_cast<float4*>(p.v_sh);
    const float4* src = reinterpret_cast<const float4*>(smem);
    for (int qd = threadIdx.x; qd < rows * gv; qd += PB) {
        const int g = qd / gv, j = qd - g * gv;
        dst[(size_t)(n0 + g) * gv + j] = (j < NV3) ? src[g * ROW + j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int DEG>
static int launch_sh_bwd(const ProjParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(SH_BWD_STAGE ? 2 : 1) * PB * 13 * 16;
    FG_CUDA(cudaFuncSetAttribute(sh_bwd_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FG_LAUNCH((sh_bwd_kernel<DEG>), ceil_div(p.N, PB), PB, smem, st, p);
    return FG_OK;
}

template <int DEG, bool VEC4>
static int launch_fwd(const ProjParams& p, cudaStream_t st) {
    using S = ShShape<(DEG >= 0 ? DEG : 0)>;
    size_t smem = DEG >= 0 ? (VEC4 ? PB * S::ROWV * 16 : PB * S::ROWF * 4) : 0;
    FG_LAUNCH((project_fwd_kernel<DEG, VEC4>), ceil_div(p.N, PB), PB, smem, st, p);
    return FG_OK;
}
template <int DEG, bool VEC4>
static int launch_bwd(const ProjParams& p, cudaStream_t st) {
    using S = ShShape<(DEG >= 0 ? DEG : 0)>;
    size_t smem = DEG >= 0 ? (VEC4 ? PB * S::ROWV * 16 : PB * S::ROWF * 4) : 0;
    FG_LAUNCH((project_bwd_kernel<DEG, VEC4>), ceil_div(p.N, PB), PB, smem, st, p);
    return FG_OK;
}

}  // namespace fg

using namespace fg;

static int check_common(int C, int N, const void* means, const void* quats, const void* scales,
                        const void* viewmats, const void* Ks, int sh_degree, int sh_bases, const void* sh) {
    FG_REQUIRE(C >= 1 && N >= 0, "C must be >= 1 and N >= 0");
    FG_REQUIRE((long long)C * N < (1ll << 31), "C*N must be < 2^31 (flatten ids are int32)");
    FG_REQUIRE(N == 0 || (means && quats && scales), "means/quats/scales must not be NULL");
    FG_REQUIRE(viewmats && Ks, "viewmats/Ks must not be NULL");
    FG_REQUIRE(sh_degree >= -1 && sh_degree <= 3, "sh_degree must be -1 (none) or 0..3");
    if (sh_degree >= 0) {
        FG_REQUIRE(sh != nullptr, "sh_coeffs must not be NULL when sh_degree >= 0");
        FG_REQUIRE(sh_bases >= (sh_degree + 1) * (sh_degree + 1), "sh_coeffs has fewer bases than sh_degree needs");
    }
    return FG_OK;
}

extern "C" int fg_project_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                              const float* viewmats, const float* Ks, int width, int height, float eps2d,
                              float near_plane, float far_plane, float radius_clip, int tile_size,
                              int sh_degree, int sh_bases, const float* sh_coeffs, const float* means_next,
                              const float* quats_next, const float* scales_next, int flow_cov, int32_t* radii,
                              float* means2d, float* depths, float* conics, float* compensations, float* feat,
                              int feat_stride, int rgb_off, int depth_off, int flow_off, float* flow_affine,
                              int32_t* tiles_per_gauss, void* stream) {
    if (C >= 1 && N == 0) return FG_OK;  // nothing to project (empty tensors have NULL data pointers)
    if (int e = check_common(C, N, means, quats, scales, viewmats, Ks, sh_degree, sh_bases, sh_coeffs)) return e;
    FG_REQUIRE(width > 0 && height > 0 && tile_size > 0, "width/height/tile_size must be positive");
    FG_REQUIRE(!flow_cov || (means_next && flow_affine), "covariance flow mode needs means_next and flow_affine");
    FG_REQUIRE(radii && means2d && depths && conics && tiles_per_gauss, "output pointers must not be NULL");
    FG_REQUIRE(feat || (sh_degree < 0 && depth_off < 0 && flow_off < 0), "feat must not be NULL");
    FG_REQUIRE(flow_off < 0 || means_next, "flow_off given without means_next");
    if (N == 0) return FG_OK;
    ProjParams p = {};
    p.C = C; p.N = N; p.means = means; p.quats = quats; p.scales = scales; p.viewmats = viewmats; p.Ks = Ks;
    p.pc = {width, height, eps2d, near_plane, far_plane, radius_clip};
    p.quats_next = quats_next; p.scales_next = scales_next; p.flow_cov = flow_cov; p.flow_affine = flow_cov ? flow_affine : nullptr;
    p.tile_size = tile_size;
    p.tile_w = (width + tile_size - 1) / tile_size;
    p.tile_h = (height + tile_size - 1) / tile_size;
    p.sh_row_floats = sh_bases * 3; p.sh = sh_coeffs; p.means_next = means_next;
    p.feat_stride = feat_stride; p.rgb_off = rgb_off; p.depth_off = depth_off; p.flow_off = flow_off;
    p.radii = radii; p.means2d = means2d; p.depths = depths; p.conics = conics; p.comps = compensations;
    p.feat = feat; p.tiles_per_gauss = tiles_per_gauss;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (p.sh_row_floats % 4 == 0) && ((uintptr_t)sh_coeffs % 16 == 0);
    switch (sh_degree) {
        case -1: return launch_fwd<-1, true>(p, st);
        case 0: return vec4 ? launch_fwd<0, true>(p, st) : launch_fwd<0, false>(p, st);
        case 1: return vec4 ? launch_fwd<1, true>(p, st) : launch_fwd<1, false>(p, st);
        case 2: return vec4 ? launch_fwd<2, true>(p, st) : launch_fwd<2, false>(p, st);
        default: return vec4 ? launch_fwd<3, true>(p, st) : launch_fwd<3, false>(p, st);
    }
}

extern "C" int fg_project_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                              const float* viewmats, const float* Ks, int width, int height, float eps2d,
                              float near_plane, float far_plane, float radius_clip, int sh_degree,
                              int sh_bases, const float* sh_coeffs, const float* means_next,
                              const float* quats_next, const float* scales_next, int flow_cov,
                              const int32_t* radii, const float* v_means2d, const float* v_depths,
                              const float* v_conics, const float* v_compensations, const float* v_feat,
                              const float* feat, int feat_stride, int rgb_off, int depth_off, int flow_off,
                              const float* v_flow_affine, float* v_means, float* v_quats, float* v_scales,
                              float* v_sh, float* v_means_next, float* v_quats_next, float* v_scales_next,
                              const fg_project_bwd_pub* pub, void* stream) {
    if (C >= 1 && N == 0) return FG_OK;
    if (int e = check_common(C, N, means, quats, scales, viewmats, Ks, sh_degree, sh_bases, sh_coeffs)) return e;
    FG_REQUIRE(!flow_cov || means_next, "covariance flow mode needs means_next");
    FG_REQUIRE(radii && v_means && v_quats && v_scales, "radii and v_means/v_quats/v_scales must not be NULL");
    FG_REQUIRE(sh_degree < 0 || v_sh || pub, "v_sh must not be NULL when sh_degree >= 0 (unless the colour gradients are published)");
    FG_REQUIRE((quats_next != nullptr) == (v_quats_next != nullptr) || !flow_cov, "v_quats_next must mirror quats_next");
    FG_REQUIRE((scales_next != nullptr) == (v_scales_next != nullptr) || !flow_cov, "v_scales_next must mirror scales_next");
    if (N == 0) return FG_OK;
    ProjParams p = {};
    p.C = C; p.N = N; p.means = means; p.quats = quats; p.scales = scales; p.viewmats = viewmats; p.Ks = Ks;
    p.pc = {width, height, eps2d, near_plane, far_plane, radius_clip};
    p.quats_next = quats_next; p.scales_next = scales_next; p.flow_cov = flow_cov;
    p.v_flow_affine = flow_cov ? v_flow_affine : nullptr;
    p.v_quats_next = flow_cov ? v_quats_next : nullptr; p.v_scales_next = flow_cov ? v_scales_next : nullptr;
    p.sh_row_floats = sh_bases * 3; p.sh = sh_coeffs; p.means_next = means_next;
    p.feat_stride = feat_stride; p.rgb_off = rgb_off; p.depth_off = depth_off; p.flow_off = flow_off;
    p.radii_in = radii; p.v_means2d = v_means2d; p.v_depths = v_depths; p.v_conics = v_conics;
    p.v_comps = v_compensations; p.v_feat = v_feat;
    p.v_means = v_means; p.v_quats = v_quats; p.v_scales = v_scales; p.v_sh = v_sh; p.v_means_next = v_means_next;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (p.sh_row_floats % 4 == 0) && ((uintptr_t)sh_coeffs % 16 == 0) && ((uintptr_t)v_sh % 16 == 0);
    if (pub) {
        FG_REQUIRE(sh_degree >= 0 && vec4 && sh_bases % 4 == 0 && feat != nullptr && v_feat != nullptr,
                   "publishing needs the streaming SH kernel: sh_degree >= 0, 16-byte rows of 4k bases, feat and v_feat");
        FG_REQUIRE(pub->campos && pub->mask && pub->prefix && pub->rgb && pub->offsets && pub->nnz && C <= PB &&
                       pub->words >= (N + 31) / 32,
                   "bad fg_project_bwd_pub");
        p.pub_campos = pub->campos; p.pub_mask = pub->mask; p.pub_prefix = pub->prefix; p.pub_rgb = pub->rgb;
        p.pub_offs = pub->offsets; p.pub_nnz = (const long long*)pub->nnz; p.pub_words = pub->words;
    }
    if (sh_degree >= 0 && vec4 && sh_bases % 4 == 0 && feat != nullptr) {
        // two kernels: streaming SH backward (writes v_sh and the direction term into v_means), then
        // the geometry VJP, which adds that term to its own v_means
        p.feat_fwd = feat;
        const int phase = pub ? pub->phase : 0;  // 1: SH kernel only, 2: geometry kernel only (after a phase-1 call)
        int e = FG_OK;
        if (phase != 2) {
            switch (sh_degree) {
                case 0: e = launch_sh_bwd<0>(p, st); break;
                case 1: e = launch_sh_bwd<1>(p, st); break;
                case 2: e = launch_sh_bwd<2>(p, st); break;
                default: e = launch_sh_bwd<3>(p, st); break;
            }
        }
        if (e || phase == 1) return e;
        p.v_mean_extra = v_means;
        p.rgb_off = -1;
        return launch_bwd<-1, true>(p, st);
    }
    switch (sh_degree) {
        case -1: return launch_bwd<-1, true>(p, st);
        case 0: return vec4 ? launch_bwd<0, true>(p, st) : launch_bwd<0, false>(p, st);
        case 1: return vec4 ? launch_bwd<1, true>(p, st) : launch_bwd<1, false>(p, st);
        case 2: return vec4 ? launch_bwd<2, true>(p, st) : launch_bwd<2, false>(p, st);
        default: return vec4 ? launch_bwd<3, true>(p, st) : launch_bwd<3, false>(p, st);
    }
}
