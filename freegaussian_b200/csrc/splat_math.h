// Per-Gaussian math of the splat-render path: quaternion/scale -> covariance, EWA
// projection with culling, real spherical harmonics, tile rectangles, and the
// hand-derived vector-Jacobian products of each.  Semantics: SURVEY.md Appendix A.2-A.4
// (gsplat 1.4.0 behaviour behind freegaussian_model.py:847-868).
//
// The functions are __host__ __device__ so that tests/host_harness can run the very
// same arithmetic on the CPU against torch.autograd through the oracle; the product
// only ever calls them from the sm_100a kernels in project.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#else
#define FG_HD inline
#endif

namespace fg {

FG_HD float rsqrt_f(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
FG_HD float fminf_(float a, float b) { return a < b ? a : b; }
FG_HD float fmaxf_(float a, float b) { return a > b ? a : b; }

// Camera: rigid world->camera [R|t] (freegaussian/utils.py:162-179) + pinhole K.
struct Camera {
    float R[9];  // row-major
    float t[3];
    float fx, fy, cx, cy;
    float pos[3];  // camera centre in world = -R^T t
};

FG_HD Camera load_camera(const float* vm /*[16] row-major 4x4*/, const float* K /*[9]*/) {
    Camera c;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c.R[3 * i + j] = vm[4 * i + j];
        c.t[i] = vm[4 * i + 3];
    }
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    for (int j = 0; j < 3; ++j)
        c.pos[j] = -(c.R[j] * c.t[0] + c.R[3 + j] * c.t[1] + c.R[6 + j] * c.t[2]);
    return c;
}

// Symmetric 3x3 stored as its upper triangle.
struct Sym3 {
    float xx, xy, xz, yy, yz, zz;
};

// ---------------------------------------------------------------- quat / covariance
// (w,x,y,z), normalised inside (Appendix A.1).
FG_HD void quat_to_rotmat(const float q[4], float R[9]) {
    float inv = rsqrt_f(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
    R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
    R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// Sigma = M M^T, M = R diag(s).
FG_HD Sym3 quat_scale_to_cov(const float q[4], const float s[3]) {
    float R[9];
    quat_to_rotmat(q, R);
    float M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * s[j];
    Sym3 c;
    c.xx = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    c.xy = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    c.xz = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    c.yy = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    c.yz = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    c.zz = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
    return c;
}

// VJP of quat_scale_to_cov.  G = dL/dSigma as a full symmetric matrix (every one of the
// nine entries treated as independent, G symmetric).  Accumulates into v_q, v_s.
FG_HD void quat_scale_to_cov_vjp(const float q[4], const float s[3], const Sym3& G, float v_q[4], float v_s[3]) {
    float R[9];
    quat_to_rotmat(q, R);
    float M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * s[j];
    const float Gm[9] = {G.xx, G.xy, G.xz, G.xy, G.yy, G.yz, G.xz, G.yz, G.zz};
    // v_M = (G + G^T) M = 2 G M
    float vM[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            vM[3 * i + j] = 2.f * (Gm[3 * i] * M[j] + Gm[3 * i + 1] * M[3 + j] + Gm[3 * i + 2] * M[6 + j]);
    float vR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) vR[3 * i + j] = vM[3 * i + j] * s[j];
    for (int j = 0; j < 3; ++j) v_s[j] += R[j] * vM[j] + R[3 + j] * vM[3 + j] + R[6 + j] * vM[6 + j];
    float inv = rsqrt_f(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    float vn[4];
    vn[0] = 2.f * (x * (vR[7] - vR[5]) + y * (vR[2] - vR[6]) + z * (vR[3] - vR[1]));
    vn[1] = 2.f * (y * (vR[1] + vR[3]) + z * (vR[2] + vR[6]) + w * (vR[7] - vR[5]) - 2.f * x * (vR[4] + vR[8]));
    vn[2] = 2.f * (x * (vR[1] + vR[3]) + w * (vR[2] - vR[6]) + z * (vR[5] + vR[7]) - 2.f * y * (vR[0] + vR[8]));
    vn[3] = 2.f * (w * (vR[3] - vR[1]) + x * (vR[2] + vR[6]) + y * (vR[5] + vR[7]) - 2.f * z * (vR[0] + vR[4]));
    float dot = vn[0] * w + vn[1] * x + vn[2] * y + vn[3] * z;
    v_q[0] += (vn[0] - dot * w) * inv;
    v_q[1] += (vn[1] - dot * x) * inv;
    v_q[2] += (vn[2] - dot * y) * inv;
    v_q[3] += (vn[3] - dot * z) * inv;
}

// ---------------------------------------------------------------- projection (A.2)
struct Projected {
    int radius;            // 0 = culled
    float mx, my;          // means2d
    float depth;           // camera z
    float ca, cb, cc;      // conic = inverse of blurred cov2d: (inv00, inv01, inv11)
    float comp;            // compensation sqrt(max(0, det_orig/det_blur))
    float a, b, c;         // blurred cov2d (needed by the covariance flow mode)
};

FG_HD void world_to_cam(const Camera& cam, const float m[3], float p[3]) {
    for (int i = 0; i < 3; ++i)
        p[i] = cam.R[3 * i] * m[0] + cam.R[3 * i + 1] * m[1] + cam.R[3 * i + 2] * m[2] + cam.t[i];
}

// Sigma_c = R Sigma R^T
FG_HD Sym3 cov_world_to_cam(const Camera& cam, const Sym3& S) {
    const float* R = cam.R;
    float A[9];  // A = R * S
    for (int i = 0; i < 3; ++i) {
        A[3 * i + 0] = R[3 * i] * S.xx + R[3 * i + 1] * S.xy + R[3 * i + 2] * S.xz;
        A[3 * i + 1] = R[3 * i] * S.xy + R[3 * i + 1] * S.yy + R[3 * i + 2] * S.yz;
        A[3 * i + 2] = R[3 * i] * S.xz + R[3 * i + 1] * S.yz + R[3 * i + 2] * S.zz;
    }
    Sym3 o;
    o.xx = A[0] * R[0] + A[1] * R[1] + A[2] * R[2];
    o.xy = A[0] * R[3] + A[1] * R[4] + A[2] * R[5];
    o.xz = A[0] * R[6] + A[1] * R[7] + A[2] * R[8];
    o.yy = A[3] * R[3] + A[4] * R[4] + A[5] * R[5];
    o.yz = A[3] * R[6] + A[4] * R[7] + A[5] * R[8];
    o.zz = A[6] * R[6] + A[7] * R[7] + A[8] * R[8];
    return o;
}

struct ProjConsts {
    int width, height;
    float eps2d, near_plane, far_plane, radius_clip;
};

// Perspective Jacobian entries with the 1.3x FoV clamp (Appendix A.2).
struct PerspJ {
    float j00, j11, j02, j12;
    float tx, ty, rz;
    bool x_in, y_in;
};
FG_HD PerspJ persp_jacobian(const Camera& cam, const float p[3], int width, int height) {
    PerspJ o;
    float x = p[0], y = p[1], z = p[2];
    float tan_fovx = 0.5f * width / cam.fx, tan_fovy = 0.5f * height / cam.fy;
    float lim_x_pos = (width - cam.cx) / cam.fx + 0.3f * tan_fovx;
    float lim_x_neg = cam.cx / cam.fx + 0.3f * tan_fovx;
    float lim_y_pos = (height - cam.cy) / cam.fy + 0.3f * tan_fovy;
    float lim_y_neg = cam.cy / cam.fy + 0.3f * tan_fovy;
    float rz = 1.f / z, rz2 = rz * rz;
    float xr = x * rz, yr = y * rz;
    o.x_in = (xr <= lim_x_pos) && (xr >= -lim_x_neg);
    o.y_in = (yr <= lim_y_pos) && (yr >= -lim_y_neg);
    o.tx = z * fminf_(lim_x_pos, fmaxf_(-lim_x_neg, xr));
    o.ty = z * fminf_(lim_y_pos, fmaxf_(-lim_y_neg, yr));
    o.rz = rz;
    o.j00 = cam.fx * rz;
    o.j11 = cam.fy * rz;
    o.j02 = -cam.fx * o.tx * rz2;
    o.j12 = -cam.fy * o.ty * rz2;
    return o;
}

// Full forward projection of one Gaussian into one camera.  Returns false (radius 0) if culled.
FG_HD bool project_gaussian(const float m[3], const Sym3& cov, const Camera& cam, const ProjConsts& pc,
                            Projected& o) {
    o.radius = 0;
    float p[3];
    world_to_cam(cam, m, p);
    if (!(p[2] >= pc.near_plane) || !(p[2] <= pc.far_plane)) return false;
    Sym3 cc = cov_world_to_cam(cam, cov);
    PerspJ J = persp_jacobian(cam, p, pc.width, pc.height);
    // cov2d = J cc J^T
    float u0 = J.j00 * cc.xx + J.j02 * cc.xz;  // (J cc) row 0, cols x,y,z
    float u1 = J.j00 * cc.xy + J.j02 * cc.yz;
    float u2 = J.j00 * cc.xz + J.j02 * cc.zz;
    float w1 = J.j11 * cc.yy + J.j12 * cc.yz;  // (J cc) row 1, cols y,z
    float w2 = J.j11 * cc.yz + J.j12 * cc.zz;
    float a0 = u0 * J.j00 + u2 * J.j02;
    float b0 = u1 * J.j11 + u2 * J.j12;
    float c0 = w1 * J.j11 + w2 * J.j12;
    float det_orig = a0 * c0 - b0 * b0;
    float a = a0 + pc.eps2d, b = b0, c = c0 + pc.eps2d;
    float det = a * c - b * b;
    if (!(det > 0.f)) return false;
    float mid = 0.5f * (a + c);
    float lam = mid + sqrtf(fmaxf_(0.01f, mid * mid - det));
    float radius = ceilf(3.f * sqrtf(lam));
    if (!(radius > pc.radius_clip)) return false;
    float mx = cam.fx * p[0] * J.rz + cam.cx;
    float my = cam.fy * p[1] * J.rz + cam.cy;
    if (mx + radius <= 0.f || mx - radius >= (float)pc.width || my + radius <= 0.f || my - radius >= (float)pc.height)
        return false;
    if (!(radius < 2.0e9f)) return false;  // inf/NaN guard: not representable as int32
    float idet = 1.f / det;
    o.radius = (int)radius;
    o.mx = mx; o.my = my; o.depth = p[2];
    o.ca = c * idet; o.cb = -b * idet; o.cc = a * idet;
    o.comp = sqrtf(fmaxf_(0.f, det_orig * idet));
    o.a = a; o.b = b; o.c = c;
    return true;
}

// Blurred 2-D covariance + projected mean with NO culling except the near plane (frame t+1 of the
// covariance flow mode, Appendix A.7).  Returns false if behind the near plane.
FG_HD bool project_cov2d(const float m[3], const Sym3& cov, const Camera& cam, const ProjConsts& pc, float& a,
                         float& b, float& c, float& u, float& v) {
    float p[3];
    world_to_cam(cam, m, p);
    if (!(p[2] >= pc.near_plane)) return false;
    Sym3 cc = cov_world_to_cam(cam, cov);
    PerspJ J = persp_jacobian(cam, p, pc.width, pc.height);
    float u0 = J.j00 * cc.xx + J.j02 * cc.xz;
    float u1 = J.j00 * cc.xy + J.j02 * cc.yz;
    float u2 = J.j00 * cc.xz + J.j02 * cc.zz;
    float w1 = J.j11 * cc.yy + J.j12 * cc.yz;
    float w2 = J.j11 * cc.yz + J.j12 * cc.zz;
    a = u0 * J.j00 + u2 * J.j02 + pc.eps2d;
    b = u1 * J.j11 + u2 * J.j12;
    c = w1 * J.j11 + w2 * J.j12 + pc.eps2d;
    u = cam.fx * p[0] * J.rz + cam.cx;
    v = cam.fy * p[1] * J.rz + cam.cy;
    return true;
}

// ---------------------------------------------------------------- covariance flow (A.7)
// A = B(t+1) B(t)^-1 - I with B the lower Cholesky factor of the blurred 2-D covariance;
// row-major (A00, A01 = 0, A10, A11).  The flow of Gaussian g at pixel p is f_g + A (p - mu_g).
FG_HD void flow_affine(const float ct[3], const float cn[3], float A[4]) {
    float l00 = sqrtf(ct[0]), l10 = ct[1] / l00, l11 = sqrtf(ct[2] - l10 * l10);
    float m00 = sqrtf(cn[0]), m10 = cn[1] / m00, m11 = sqrtf(cn[2] - m10 * m10);
    float i00 = 1.f / l00, i11 = 1.f / l11;
    A[0] = m00 * i00 - 1.f;
    A[1] = 0.f;
    A[2] = m10 * i00 - m11 * l10 * i00 * i11;
    A[3] = m11 * i11 - 1.f;
}
// VJP: vA[4] -> v_ct[3], v_cn[3] (gradients w.r.t. the stored (a,b,c) of both covariances; accumulated)
FG_HD void flow_affine_vjp(const float ct[3], const float cn[3], const float vA[4], float v_ct[3], float v_cn[3]) {
    float l00 = sqrtf(ct[0]), l10 = ct[1] / l00, l11 = sqrtf(ct[2] - l10 * l10);
    float m00 = sqrtf(cn[0]), m10 = cn[1] / m00, m11 = sqrtf(cn[2] - m10 * m10);
    float i00 = 1.f / l00, i11 = 1.f / l11;
    float v_m00 = vA[0] * i00;
    float v_i00 = vA[0] * m00 + vA[2] * (m10 - m11 * l10 * i11);
    float v_m10 = vA[2] * i00;
    float v_m11 = -vA[2] * l10 * i00 * i11 + vA[3] * i11;
    float v_l10 = -vA[2] * m11 * i00 * i11;
    float v_i11 = -vA[2] * m11 * l10 * i00 + vA[3] * m11;
    float v_l00 = -v_i00 * i00 * i00;
    float v_l11 = -v_i11 * i11 * i11;
    // Cholesky VJP, frame t: l00 = sqrt(a), l10 = b/l00, l11 = sqrt(c - l10^2)
    v_ct[2] += v_l11 / (2.f * l11);
    v_l10 += -v_l11 * l10 / l11;
    v_ct[1] += v_l10 * i00;
    v_l00 += -v_l10 * l10 * i00;
    v_ct[0] += v_l00 / (2.f * l00);
    // frame t+1
    v_cn[2] += v_m11 / (2.f * m11);
    v_m10 += -v_m11 * m10 / m11;
    v_cn[1] += v_m10 / m00;
    v_m00 += -v_m10 * m10 / m00;
    v_cn[0] += v_m00 / (2.f * m00);
}

// Project a point only (frame t+1 mean for the flow channel, Appendix A.7).
FG_HD bool project_point(const float m[3], const Camera& cam, float near_plane, float& u, float& v) {
    float p[3];
    world_to_cam(cam, m, p);
    if (!(p[2] >= near_plane)) return false;
    float rz = 1.f / p[2];
    u = cam.fx * p[0] * rz + cam.cx;
    v = cam.fy * p[1] * rz + cam.cy;
    return true;
}
FG_HD void project_point_vjp(const float m[3], const Camera& cam, float v_u, float v_v, float v_m[3]) {
    float p[3];
    world_to_cam(cam, m, p);
    float rz = 1.f / p[2], rz2 = rz * rz;
    float vp[3] = {cam.fx * rz * v_u, cam.fy * rz * v_v, -(cam.fx * p[0] * v_u + cam.fy * p[1] * v_v) * rz2};
    for (int j = 0; j < 3; ++j) v_m[j] += cam.R[j] * vp[0] + cam.R[3 + j] * vp[1] + cam.R[6 + j] * vp[2];
}

// VJP of project_gaussian for a non-culled Gaussian.
//   v_m2d[2], v_depth, v_conic[3] (d/d(stored A,B,C)), v_comp, v_cov2d[3] (d/d(stored blurred a,b,c); may be null)
//   ->  accumulates v_mean[3], G (dL/dSigma, full-symmetric)
FG_HD void project_gaussian_vjp(const float m[3], const Sym3& cov, const Camera& cam, const ProjConsts& pc,
                                const float v_m2d[2], float v_depth, const float v_conic[3], float v_comp,
                                float v_mean[3], Sym3& G, const float* v_cov2d = nullptr) {
    float p[3];
    world_to_cam(cam, m, p);
    Sym3 cc = cov_world_to_cam(cam, cov);
    PerspJ J = persp_jacobian(cam, p, pc.width, pc.height);
    float u0 = J.j00 * cc.xx + J.j02 * cc.xz;
    float u1 = J.j00 * cc.xy + J.j02 * cc.yz;
    float u2 = J.j00 * cc.xz + J.j02 * cc.zz;
    float w0 = J.j11 * cc.xy + J.j12 * cc.xz;
    float w1 = J.j11 * cc.yy + J.j12 * cc.yz;
    float w2 = J.j11 * cc.yz + J.j12 * cc.zz;
    float a0 = u0 * J.j00 + u2 * J.j02;
    float b0 = u1 * J.j11 + u2 * J.j12;
    float c0 = w1 * J.j11 + w2 * J.j12;
    float a = a0 + pc.eps2d, b = b0, c = c0 + pc.eps2d;
    float det = a * c - b * b;
    float idet = 1.f / det;
    float A = c * idet, B = -b * idet, C = a * idet;  // conic
    // V = dL/dcov2d (full symmetric) = -Ci Vc Ci, Vc = [[vA, vB/2],[vB/2, vC]]
    float hB = 0.5f * v_conic[1];
    float X00 = v_conic[0] * A + hB * B, X01 = v_conic[0] * B + hB * C;
    float X10 = hB * A + v_conic[2] * B, X11 = hB * B + v_conic[2] * C;
    float V00 = -(A * X00 + B * X10);
    float V01 = -(A * X01 + B * X11);
    float V11 = -(B * X01 + C * X11);
    if (v_cov2d) {  // stored b stands for both off-diagonal entries of the symmetric matrix
        V00 += v_cov2d[0];
        V01 += 0.5f * v_cov2d[1];
        V11 += v_cov2d[2];
    }
    if (v_comp != 0.f) {
        // comp = sqrt(max(0, det_orig/det)); d(comp^2)/dcov2d = (1-comp^2) Ci - eps2d det(Ci) I
        float det_orig = a0 * c0 - b0 * b0;
        float comp = sqrtf(fmaxf_(0.f, det_orig * idet));
        if (comp > 0.f) {
            float vs = v_comp * 0.5f / comp;
            float om = 1.f - comp * comp;
            float dci = A * C - B * B;
            V00 += vs * (om * A - pc.eps2d * dci);
            V01 += vs * (om * B);
            V11 += vs * (om * C - pc.eps2d * dci);
        }
    }
    // v_cc = J^T V J  (3x3 symmetric)
    // rows of (V J): r0 = (V00 j00, V01 j11, V00 j02 + V01 j12), r1 = (V01 j00, V11 j11, V01 j02 + V11 j12)
    float r00 = V00 * J.j00, r01 = V01 * J.j11, r02 = V00 * J.j02 + V01 * J.j12;
    float r10 = V01 * J.j00, r11 = V11 * J.j11, r12 = V01 * J.j02 + V11 * J.j12;
    Sym3 vcc;
    vcc.xx = J.j00 * r00;
    vcc.xy = J.j00 * r01;
    vcc.xz = J.j00 * r02;
    vcc.yy = J.j11 * r11;
    vcc.yz = J.j11 * r12;
    vcc.zz = J.j02 * r02 + J.j12 * r12;
    // v_J = 2 V (J cc): (J cc) rows are (u0,u1,u2), (w0,w1,w2)
    float vJ00 = 2.f * (V00 * u0 + V01 * w0);
    float vJ02 = 2.f * (V00 * u2 + V01 * w2);
    float vJ11 = 2.f * (V01 * u1 + V11 * w1);
    float vJ12 = 2.f * (V01 * u2 + V11 * w2);
    // v_p (camera-space mean)
    float rz = J.rz, rz2 = rz * rz, rz3 = rz2 * rz;
    float vp[3];
    vp[0] = cam.fx * rz * v_m2d[0];
    vp[1] = cam.fy * rz * v_m2d[1];
    vp[2] = -(cam.fx * p[0] * v_m2d[0] + cam.fy * p[1] * v_m2d[1]) * rz2 + v_depth;
    if (J.x_in) vp[0] += -cam.fx * rz2 * vJ02; else vp[2] += -cam.fx * rz3 * vJ02 * J.tx;
    if (J.y_in) vp[1] += -cam.fy * rz2 * vJ12; else vp[2] += -cam.fy * rz3 * vJ12 * J.ty;
    vp[2] += -cam.fx * rz2 * vJ00 - cam.fy * rz2 * vJ11 + 2.f * cam.fx * J.tx * rz3 * vJ02 +
             2.f * cam.fy * J.ty * rz3 * vJ12;
    const float* R = cam.R;
    for (int j = 0; j < 3; ++j) v_mean[j] += R[j] * vp[0] + R[3 + j] * vp[1] + R[6 + j] * vp[2];
    // G += R^T vcc R
    float Bm[9];  // vcc * R
    const float S[9] = {vcc.xx, vcc.xy, vcc.xz, vcc.xy, vcc.yy, vcc.yz, vcc.xz, vcc.yz, vcc.zz};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Bm[3 * i + j] = S[3 * i] * R[j] + S[3 * i + 1] * R[3 + j] + S[3 * i + 2] * R[6 + j];
    G.xx += R[0] * Bm[0] + R[3] * Bm[3] + R[6] * Bm[6];
    G.xy += R[0] * Bm[1] + R[3] * Bm[4] + R[6] * Bm[7];
    G.xz += R[0] * Bm[2] + R[3] * Bm[5] + R[6] * Bm[8];
    G.yy += R[1] * Bm[1] + R[4] * Bm[4] + R[7] * Bm[7];
    G.yz += R[1] * Bm[2] + R[4] * Bm[5] + R[7] * Bm[8];
    G.zz += R[2] * Bm[2] + R[5] * Bm[5] + R[8] * Bm[8];
}

// ---------------------------------------------------------------- spherical harmonics (A.3)
// basis values for a *normalised* direction; writes (degree+1)^2 entries.
FG_HD void sh_basis(int degree, float x, float y, float z, float* B) {
    B[0] = 0.2820947917738781f;
    if (degree < 1) return;
    B[1] = -0.48860251190292f * y;
    B[2] = 0.48860251190292f * z;
    B[3] = -0.48860251190292f * x;
    if (degree < 2) return;
    float z2 = z * z;
    float fTmp0B = -1.092548430592079f * z;
    float fC1 = x * x - y * y;
    float fS1 = 2.f * x * y;
    B[4] = 0.5462742152960395f * fS1;
    B[5] = fTmp0B * y;
    B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    B[7] = fTmp0B * x;
    B[8] = 0.5462742152960395f * fC1;
    if (degree < 3) return;
    float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    float fTmp1B = 1.445305721320277f * z;
    float fC2 = x * fC1 - y * fS1;
    float fS2 = x * fS1 + y * fC1;
    B[9] = -0.5900435899266435f * fS2;
    B[10] = fTmp1B * fS1;
    B[11] = fTmp0C * y;
    B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    B[13] = fTmp0C * x;
    B[14] = fTmp1B * fC1;
    B[15] = -0.5900435899266435f * fC2;
}

// d(sum_k s_k B_k)/d(x,y,z) for a normalised direction, s_k = coeff_k . v_rgb
FG_HD void sh_basis_vjp(int degree, float x, float y, float z, const float* s, float vd[3]) {
    vd[0] = vd[1] = vd[2] = 0.f;
    if (degree < 1) return;
    const float C1 = 0.48860251190292f;
    vd[1] += -C1 * s[1];
    vd[2] += C1 * s[2];
    vd[0] += -C1 * s[3];
    if (degree < 2) return;
    const float k2 = 0.5462742152960395f, k1 = 1.092548430592079f, k6 = 0.9461746957575601f;
    vd[0] += 2.f * k2 * y * s[4];
    vd[1] += 2.f * k2 * x * s[4];
    vd[1] += -k1 * z * s[5];
    vd[2] += -k1 * y * s[5];
    vd[2] += 2.f * k6 * z * s[6];
    vd[0] += -k1 * z * s[7];
    vd[2] += -k1 * x * s[7];
    vd[0] += 2.f * k2 * x * s[8];
    vd[1] += -2.f * k2 * y * s[8];
    if (degree < 3) return;
    const float k9 = 0.5900435899266435f, k10 = 1.445305721320277f, k11 = 2.285228997322329f;
    float z2 = z * z;
    float fTmp0C = -k11 * z2 + 0.4570457994644658f;
    float xx_yy = x * x - y * y;
    vd[0] += -k9 * 6.f * x * y * s[9];
    vd[1] += -k9 * 3.f * xx_yy * s[9];
    vd[0] += k10 * 2.f * y * z * s[10];
    vd[1] += k10 * 2.f * x * z * s[10];
    vd[2] += k10 * 2.f * x * y * s[10];
    vd[1] += fTmp0C * s[11];
    vd[2] += -2.f * k11 * z * y * s[11];
    vd[2] += (3.f * 1.865881662950577f * z2 - 1.119528997770346f) * s[12];
    vd[0] += fTmp0C * s[13];
    vd[2] += -2.f * k11 * z * x * s[13];
    vd[0] += k10 * 2.f * x * z * s[14];
    vd[1] += -k10 * 2.f * y * z * s[14];
    vd[2] += k10 * xx_yy * s[14];
    vd[0] += -k9 * 3.f * xx_yy * s[15];
    vd[1] += k9 * 6.f * x * y * s[15];
}

// ---------------------------------------------------------------- tiles (A.4)
struct TileRect {
    int x0, x1, y0, y1;  // [x0,x1) x [y0,y1)
};
FG_HD TileRect tile_rect(float mx, float my, int radius, int tile_size, int tile_w, int tile_h) {
    float ts = (float)tile_size;
    float tr, tx, ty;
    if ((tile_size & (tile_size - 1)) == 0) {  // power of two (16 in every caller): x * 2^-k == x / 2^k exactly, bit for bit
        const float inv = 1.0f / ts;
        tr = (float)radius * inv; tx = mx * inv; ty = my * inv;
    } else {
        tr = (float)radius / ts; tx = mx / ts; ty = my / ts;
    }
    float fx0 = floorf(tx - tr), fx1 = ceilf(tx + tr), fy0 = floorf(ty - tr), fy1 = ceilf(ty + tr);
    TileRect r;
    r.x0 = (int)fminf_(fmaxf_(fx0, 0.f), (float)tile_w);
    r.x1 = (int)fminf_(fmaxf_(fx1, 0.f), (float)tile_w);
    r.y0 = (int)fminf_(fmaxf_(fy0, 0.f), (float)tile_h);
    r.y1 = (int)fminf_(fmaxf_(fy1, 0.f), (float)tile_h);
    return r;
}

FG_HD int tile_bits_of(int n_tiles) {  // floor(log2(n)) + 1
    int b = 0;
    while (n_tiles > 0) { ++b; n_tiles >>= 1; }
    return b;
}

}  // namespace fg
