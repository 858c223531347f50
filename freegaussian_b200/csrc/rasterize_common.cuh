// Shared pieces of the compositing kernels: tile/pixel mapping,
// the shared-memory Gaussian batch, and the one alpha formula forward and backward share.
#pragma once
#include "common.cuh"

namespace fg {

constexpr int TILE = 16;              // pixels per tile side (freegaussian_model.py:806)
constexpr int TILE_PIX = TILE * TILE;  // 256 threads, one pixel each
constexpr int BATCH = 256;            // Gaussians staged in shared memory per round
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float ALPHA_MAX = 0.999f;
constexpr float T_STOP = 1e-4f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// Warp w owns an 8x4 pixel patch of the tile (2 patches across, 4 down): spatially compact
// warps make the per-warp "nobody touches this Gaussian" vote succeed far more often than
// 32x1 or 16x2 strips do.
__device__ __forceinline__ void tile_pixel(int tid, int& lx, int& ly) {
    const int warp = tid >> 5, lane = tid & 31;
    lx = ((warp & 1) << 3) + (lane & 7);
    ly = ((warp >> 1) << 2) + (lane >> 3);
}

// Geometry of one staged Gaussian.  The conic is pre-scaled by 0.5 log2(e),
//   a1 = 0.5 log2e A,  b1 = 0.5 log2e B,  c1 = 0.5 log2e C,
// so that with u = a1 dx + b1 dy and v = b1 dx + c1 dy the exponent is p = u dx + v dy = log2e sigma
// and alpha = opac 2^-p.  u and v are also what the backward pass needs: d sigma / d(dx, dy) = 2 ln2 (u, v).
struct GeomA {
    float x, y, opac, a1;
};
struct GeomB {
    float b1, c1;
    int id;  // flatten id c*N+n
    float pad;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// alpha for pixel (px,py).  Returns false when the Gaussian is skipped (sigma < 0 or alpha < 1/255).
// vis = exp(-sigma), raw = opac * vis (alpha before the 0.999 clamp).
__device__ __forceinline__ bool eval_alpha(const GeomA& a, const GeomB& b, float px, float py, float& dx,
                                           float& dy, float& u, float& v, float& vis, float& raw, float& alpha) {
    dx = a.x - px;
    dy = a.y - py;
    u = fmaf(a.a1, dx, b.b1 * dy);
    v = fmaf(b.c1, dy, b.b1 * dx);
    const float p = fmaf(u, dx, v * dy);
    vis = ex2_approx(-p);
    raw = a.opac * vis;
    alpha = fminf(ALPHA_MAX, raw);
    return (p >= 0.f) && (alpha >= ALPHA_MIN);
}

// ---- packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2): both pixels of a thread in one instruction --------------------
// The compositing kernels are issue-bound (ncu r1: issue slots 76-85 % busy, FMA pipe ~40 %): with two pixels per thread
// (four rows apart, the two 8x4 patches of the culling mask) the per-pixel arithmetic runs on float2 = (pixel 0, pixel 1)
// operands.  A scalar operand is broadcast by the instruction itself (`R.F32` operand form in SASS), so per-Gaussian
// values need no duplication.  tools/ubench/f32x2.cu: FFMA2 sustains the FFMA flop rate with half the issue slots.
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

// alpha of both pixels; the same operations in the same order as eval_alpha, so the one- and two-pixel kernels and the
// forward and backward passes agree bit for bit.  npy = (-py0, -py1).
__device__ __forceinline__ void eval_alpha_pair(const GeomA& a, const GeomB& b, float px, float2 npy, float& dx, float2& dy,
                                                float2& u, float2& v, float2& vis, float2& raw, float2& alpha, bool& ok0,
                                                bool& ok1) {
    dx = a.x - px;
    dy = __fadd2_rn(bc2(a.y), npy);
    u = __ffma2_rn(bc2(a.a1), bc2(dx), __fmul2_rn(bc2(b.b1), dy));
    v = __ffma2_rn(bc2(b.c1), dy, bc2(b.b1 * dx));
    const float2 pw = __ffma2_rn(u, bc2(dx), __fmul2_rn(v, dy));
    vis = make_float2(ex2_approx(-pw.x), ex2_approx(-pw.y));
    raw = __fmul2_rn(bc2(a.opac), vis);
    alpha = make_float2(fminf(ALPHA_MAX, raw.x), fminf(ALPHA_MAX, raw.y));
    ok0 = (pw.x >= 0.f) && (alpha.x >= ALPHA_MIN);
    ok1 = (pw.y >= 0.f) && (alpha.y >= ALPHA_MIN);
}

// ---- shared memory by explicit 32-bit address --------------------------------------------------
// The staged Gaussian records live at fixed offsets from one base (slot t at base + 16 t + k * 4096),
// so the hot loops form ONE address per Gaussian and use immediate offsets; through C++ arrays the
// compiler re-derived every array's base (via SR_CgaCtaId) inside the loop (ncu r1z: ~11 of 145
// instructions per step in the backward kernel).
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ float2 lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "f"(v));
}
constexpr int REC_STRIDE = BATCH * 16;  // bytes between the float4 arrays of the staged records

// ---- per-warp culling ---------------------------------------------------------------------
// Which of the tile's eight 8x4 pixel patches (= warps) can a Gaussian touch at all, i.e. has a
// pixel with alpha >= 1/255?  alpha >= 1/255  <=>  power <= log2(255 opac) (pre-scaled units), so
// a patch is dropped when a lower bound of the power over the patch's rectangle of pixel
// centres exceeds that threshold.  The bound is the exact minimum of the convex quadratic over
// the continuous rectangle (zero if the mean lies inside, else the best of the four edge
// minima), with a small safety margin: conservative, so compositing results are unchanged.
// ncu r1g: 60 % of the (warp, Gaussian) steps had no valid pixel and cost ~25 instructions each.
// minimise q_cc c^2 + q_cv c v + q_vv v^2 over v in [vlo, vhi]; nh_cv_over_vv = -0.5 q_cv / q_vv
__device__ __forceinline__ float edge_min(float q_cc, float q_cv, float q_vv, float nh_cv_over_vv, float c, float vlo,
                                          float vhi) {
    const float v = fminf(fmaxf(nh_cv_over_vv * c, vlo), vhi);
    return fmaf(q_cc * c, c, v * fmaf(q_cv, c, q_vv * v));
}

__device__ __forceinline__ unsigned patch_mask(float mx, float my, float opac, float qa, float qb, float qc,
                                               float tile_x0, float tile_y0) {
    // tile_x0/tile_y0: centre of the tile's first pixel
    const float thr = __log2f(255.f * opac) * 1.00002f + 2e-4f;
    if (!(thr > 0.f) || !(qa > 0.f) || !(qc > 0.f)) return (thr > 0.f) ? 0xffu : 0u;  // degenerate conic: keep
    const float kx = -0.5f * qb / qa, ky = -0.5f * qb / qc;  // argmin slopes, shared by all patches
    // the eight patches share two x-intervals and four y-intervals of delta = mu - p
    float dxlo[2], dxhi[2], dylo[4], dyhi[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) { dxhi[i] = mx - (tile_x0 + 8.f * i); dxlo[i] = dxhi[i] - 7.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { dyhi[j] = my - (tile_y0 + 4.f * j); dylo[j] = dyhi[j] - 3.f; }
    unsigned mask = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int i = w & 1, j = w >> 1;
        float pmin;
        if (dxlo[i] <= 0.f && dxhi[i] >= 0.f && dylo[j] <= 0.f && dyhi[j] >= 0.f) {
            pmin = 0.f;
        } else {
            pmin = fminf(fminf(edge_min(qa, qb, qc, ky, dxlo[i], dylo[j], dyhi[j]),
                               edge_min(qa, qb, qc, ky, dxhi[i], dylo[j], dyhi[j])),
                         fminf(edge_min(qc, qb, qa, kx, dylo[j], dxlo[i], dxhi[i]),
                               edge_min(qc, qb, qa, kx, dyhi[j], dxlo[i], dxhi[i])));
        }
        if (pmin <= thr) mask |= 1u << w;
    }
    return mask;
}

// Compact, in order, the batch slots t in [t_min, bs) whose patch mask has this warp's bit set.
__device__ __forceinline__ int build_warp_list(const unsigned char* sMask, unsigned char* list, int warp, int lane,
                                               int t_min, int bs) {
    int n = 0;
    const unsigned lt = (1u << lane) - 1;
#pragma unroll
    for (int c = 0; c < BATCH / 32; ++c) {
        const int t = c * 32 + lane;
        const bool keep = t >= t_min && t < bs && ((sMask[t] >> warp) & 1u);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) list[n + __popc(m & lt)] = (unsigned char)t;
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

}  // namespace fg
