// Fused Adam over the Gaussian parameter groups.
//
// The reference steps eight torch.optim.Adam instances, one per parameter group
// (freegaussian_config.py:48-90: eps = 1e-15, default betas, no weight decay), i.e. ~6 kernels per
// group per step.  Here every group is a "segment" of one launch; features_dc / features_rest,
// which the reference concatenates into the [N,16,3] SH tensor every step
// (freegaussian_model.py:801), can be ONE segment with a per-column learning rate, so the
// concatenation and its backward split disappear.  HBM-bound: 16 B read + 12 B written per element.
//
// Arithmetic follows torch.optim.Adam's single-tensor path operation by operation:
//   m = m + (1-b1) (g - m);  v = b2 v + (1-b2) g g;
//   p = p - (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#include "common.cuh"

namespace fg {
namespace {

constexpr int AB = 256;          // threads per block
constexpr int A_ITEMS = 4;       // float4 per thread
constexpr int A_TILE = AB * A_ITEMS * 4;  // elements per block

struct AdamSegDev {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    long long first;  // index of p[0] inside the full tensor (column phase of a shard)
    int row_len, split;
    float step0, step1;  // lr / (1 - b1^t) for columns < split and >= split
    int vec;
};

struct AdamParams {
    AdamSegDev seg[FG_ADAM_MAX_SEGMENTS];
    int blk_end[FG_ADAM_MAX_SEGMENTS];  // exclusive prefix of blocks per segment
    int n_seg;
    float b2, omb1, omb2, eps, inv_bc2_sqrt;  // beta2, 1-beta1, 1-beta2 (rounded from double like torch's scalars)
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float step, const AdamParams& P) {
    m = m + P.omb1 * (g - m);
    v = v * P.b2 + P.omb2 * g * g;
    const float denom = sqrtf(v) * P.inv_bc2_sqrt + P.eps;
    p = p - step * (m / denom);
}

__global__ void __launch_bounds__(AB) adam_kernel(const __grid_constant__ AdamParams P) {
    pdl_wait();
    int s = 0;
    while (s < P.n_seg - 1 && (int)blockIdx.x >= P.blk_end[s]) ++s;
    const AdamSegDev& S = P.seg[s];
    const int blk = blockIdx.x - (s ? P.blk_end[s - 1] : 0);
    const long long base = (long long)blk * A_TILE;
    const bool two = S.split > 0 && S.split < S.row_len;
    if (S.vec) {
#pragma unroll
        for (int it = 0; it < A_ITEMS; ++it) {
            const long long e = base + ((long long)it * AB + threadIdx.x) * 4;
            if (e >= S.n) break;
            if (e + 4 <= S.n) {
                float4 p = *reinterpret_cast<float4*>(S.p + e);
                const float4 g = __ldg(reinterpret_cast<const float4*>(S.g + e));
                float4 m = *reinterpret_cast<float4*>(S.m + e);
                float4 v = *reinterpret_cast<float4*>(S.v + e);
                float st[4] = {S.step0, S.step0, S.step0, S.step0};
                if (two) {
                    const unsigned c0 = (unsigned)((S.first + e) % S.row_len);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        unsigned c = c0 + j;
                        if (c >= (unsigned)S.row_len) c -= S.row_len;
                        st[j] = c < (unsigned)S.split ? S.step0 : S.step1;
                    }
                }
                adam1(p.x, g.x, m.x, v.x, st[0], P);
                adam1(p.y, g.y, m.y, v.y, st[1], P);
                adam1(p.z, g.z, m.z, v.z, st[2], P);
                adam1(p.w, g.w, m.w, v.w, st[3], P);
                *reinterpret_cast<float4*>(S.p + e) = p;
                *reinterpret_cast<float4*>(S.m + e) = m;
                *reinterpret_cast<float4*>(S.v + e) = v;
            } else {
                for (long long q = e; q < S.n; ++q) {
                    const float st = (two && (int)((S.first + q) % S.row_len) >= S.split) ? S.step1 : S.step0;
                    float p = S.p[q], m = S.m[q], v = S.v[q];
                    adam1(p, S.g[q], m, v, st, P);
                    S.p[q] = p; S.m[q] = m; S.v[q] = v;
                }
            }
        }
    } else {
        for (int it = 0; it < A_ITEMS * 4; ++it) {
            const long long q = base + (long long)it * AB + threadIdx.x;
            if (q >= S.n) break;
            const float st = (two && (int)((S.first + q) % S.row_len) >= S.split) ? S.step1 : S.step0;
            float p = S.p[q], m = S.m[q], v = S.v[q];
            adam1(p, S.g[q], m, v, st, P);
            S.p[q] = p; S.m[q] = m; S.v[q] = v;
        }
    }
}

}  // namespace
}  // namespace fg

using namespace fg;

extern "C" int fg_adam_step(int n_segments, const fg_adam_segment* segments, int step, double beta1, double beta2,
                            double eps, void* stream) {
    FG_REQUIRE(n_segments >= 0 && n_segments <= FG_ADAM_MAX_SEGMENTS, "segment count");
    FG_REQUIRE(step >= 1, "Adam step counts from 1");
    FG_REQUIRE(n_segments == 0 || segments, "NULL pointer");
    AdamParams P{};
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    P.b2 = (float)beta2; P.omb1 = (float)(1.0 - beta1); P.omb2 = (float)(1.0 - beta2); P.eps = (float)eps;
    P.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    int blocks = 0, k = 0;
    for (int i = 0; i < n_segments; ++i) {
        const fg_adam_segment& a = segments[i];
        FG_REQUIRE(a.n >= 0 && a.first >= 0 && a.row_len >= 0 && a.split >= 0, "segment shape");
        if (a.n == 0) continue;
        FG_REQUIRE(a.param && a.grad && a.exp_avg && a.exp_avg_sq, "NULL pointer");
        AdamSegDev& S = P.seg[k];
        S.p = a.param; S.g = a.grad; S.m = a.exp_avg; S.v = a.exp_avg_sq;
        S.n = a.n; S.first = a.first;
        S.row_len = a.row_len > 0 ? a.row_len : 1;
        S.split = a.row_len > 0 ? a.split : 0;
        S.step0 = (float)((double)a.lr / bc1);
        S.step1 = (float)((double)a.lr_rest / bc1);
        S.vec = (((uintptr_t)a.param | (uintptr_t)a.grad | (uintptr_t)a.exp_avg | (uintptr_t)a.exp_avg_sq) % 16) == 0;
        const long long nb = (a.n + A_TILE - 1) / A_TILE;
        FG_REQUIRE(blocks + nb < (1ll << 31), "too many elements");
        blocks += (int)nb;
        P.blk_end[k] = blocks;
        ++k;
    }
    P.n_seg = k;
    if (blocks == 0) return FG_OK;
    FG_LAUNCH(adam_kernel, blocks, AB, 0, stream, P);
    return FG_OK;
}
