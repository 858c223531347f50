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

// Geometry of one staged Gaussian.  Conic is pre-scaled so that
//   alpha = opac * 2^-(qa dx^2 + qb dx dy + qc dy^2),  qa = 0.5 log2e A, qb = log2e B, qc = 0.5 log2e C.
struct GeomA {
    float x, y, opac, qa;
};
struct GeomB {
    float qb, qc;
    int id;  // flatten id c*N+n
    float pad;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// power (scaled sigma) and alpha for pixel (px,py).  Returns false when the Gaussian is
// skipped (sigma < 0 or alpha < 1/255).  vis = exp(-sigma).
__device__ __forceinline__ bool eval_alpha(const GeomA& a, const GeomB& b, float px, float py, float& dx,
                                           float& dy, float& vis, float& alpha) {
    dx = a.x - px;
    dy = a.y - py;
    float p = fmaf(a.qa * dx, dx, fmaf(b.qc * dy, dy, b.qb * dx * dy));
    vis = ex2_approx(-p);
    alpha = fminf(ALPHA_MAX, a.opac * vis);
    return (p >= 0.f) && (alpha >= ALPHA_MIN);
}

}  // namespace fg
