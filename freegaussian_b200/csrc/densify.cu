// Refinement (split / duplicate / cull with Adam-state surgery) as plan -> map -> gather kernels
// (freegaussian_model.py:404-571, :313-367).  The reference builds the new
// parameter tensors with ~40 boolean-mask indexing ops, torch.cat per group and per Adam moment,
// and several .item() syncs; here the masks are one kernel, the positions one scan, and every
// array moves once through a row gather.  Output order is the reference's: kept originals,
// children sample-major (`repeat(samps, 1)`), duplicates.
#include "common.cuh"

namespace fg {
namespace {

constexpr int RB = 256;

struct PlanParams {
    long long N;
    const float* scales;
    const float* opac;
    const float* gn;
    const float* vc;
    const float* ms;
    fg_refine_config c;
    int* flags;  // [4N]: keep_orig | keep_child | keep_dup | split
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(RB) refine_flags_kernel(PlanParams p) {
    pdl_wait();
    const long long n = (long long)blockIdx.x * RB + threadIdx.x;
    if (n >= p.N) return;
    const float s0 = p.scales[3 * n], s1 = p.scales[3 * n + 1], s2 = p.scales[3 * n + 2];
    const float smax = fmaxf(fmaxf(expf(s0), expf(s1)), expf(s2));
    const float msz = p.ms ? p.ms[n] : 0.f;
    const bool screen = p.c.use_screen && p.ms;
    // children (and, because split_gaussians rescales its parents in place at :536 before the dup mask
    // is taken at :430, duplicates of split parents) carry scale' = log(exp(scale) / 1.6)
    const float c0 = expf(logf(expf(s0) / 1.6f)), c1 = expf(logf(expf(s1) / 1.6f)), c2 = expf(logf(expf(s2) / 1.6f));
    const float cmax = fmaxf(fmaxf(c0, c1), c2);
    bool split = false, dup = false;
    if (p.c.densify) {
        const float avg = __fdiv_rn(p.gn[n], p.vc[n]) * 0.5f * p.c.max_dim;
        const bool high = avg > p.c.densify_grad_thresh;
        split = (smax > p.c.densify_size_thresh) && high;
        if (screen) split |= msz > p.c.split_screen_size;
        dup = ((split ? cmax : smax) <= p.c.densify_size_thresh) && high;
    }
    const bool transparent = sigmoidf_(p.opac[n]) < p.c.cull_alpha_thresh;
    bool big = false, big_child = false, big_dup = false;
    if (p.c.cull_big) {
        big = smax > p.c.cull_scale_thresh || (screen && msz > p.c.cull_screen_size);
        big_child = cmax > p.c.cull_scale_thresh;  // new rows carry max_size 0
        big_dup = (split ? cmax : smax) > p.c.cull_scale_thresh;
    }
    p.flags[n] = !(transparent || split || big);
    p.flags[p.N + n] = split && !(transparent || big_child);
    p.flags[2 * p.N + n] = dup && !(transparent || big_dup);
    p.flags[3 * p.N + n] = split;
}

__global__ void refine_counts_kernel(long long N, int* plan, const long long* total, long long* counts) {
    pdl_wait();
    const int t = (int)*total;
    plan[4 * N] = t;
    counts[0] = plan[N];
    counts[1] = plan[2 * N] - plan[N];
    counts[2] = plan[3 * N] - plan[2 * N];
    counts[3] = t - plan[3 * N];
}

__global__ void __launch_bounds__(RB) refine_map_kernel(long long N, const int* __restrict__ plan,
                                                        const long long* __restrict__ counts, int samps,
                                                        long long n_out, int* __restrict__ src,
                                                        int* __restrict__ sample_row) {
    pdl_wait();
    const long long n = (long long)blockIdx.x * RB + threadIdx.x;
    if (n >= N) return;
    const long long n_ko = counts[0], n_kc = counts[1], n_split = counts[3];
    if (plan[n + 1] != plan[n]) {
        src[plan[n]] = (int)n;
        sample_row[plan[n]] = -1;
    }
    if (plan[N + n + 1] != plan[N + n]) {
        const long long rank_kc = plan[N + n] - n_ko;
        const long long rank_split = plan[3 * N + n] - plan[3 * N];
        for (int k = 0; k < samps; ++k) {
            const long long d = n_ko + (long long)k * n_kc + rank_kc;
            if (d < n_out) {
                src[d] = (int)n;
                sample_row[d] = (int)((long long)k * n_split + rank_split);
            }
        }
    }
    if (plan[2 * N + n + 1] != plan[2 * N + n]) {
        const long long d = n_ko + (long long)samps * n_kc + (plan[2 * N + n] - plan[2 * N]);
        if (d < n_out) {
            src[d] = (int)n;
            sample_row[d] = (plan[3 * N + n + 1] != plan[3 * N + n]) ? -2 : -1;  // -2: copy of a rescaled parent
        }
    }
}

struct GatherParams {
    fg_refine_array a[FG_REFINE_MAX_ARRAYS];
    long long n_out, n_keep;
    const int* src;
};

__global__ void __launch_bounds__(RB) refine_gather_kernel(const __grid_constant__ GatherParams p) {
    pdl_wait();
    const fg_refine_array& A = p.a[blockIdx.y];
    const long long total = p.n_out * A.row_floats;
    for (long long e = (long long)blockIdx.x * RB + threadIdx.x; e < total; e += (long long)gridDim.x * RB) {
        const long long d = e / A.row_floats;
        const int c = (int)(e - d * A.row_floats);
        float v = 0.f;
        if (!(A.zero_new && d >= p.n_keep)) v = __ldg(A.in + (long long)p.src[d] * A.row_floats + c);
        A.out[e] = v;
    }
}

__global__ void __launch_bounds__(RB) refine_children_kernel(long long n_keep, long long n_children,
                                                             const int* __restrict__ src,
                                                             const int* __restrict__ sample_row,
                                                             const float* __restrict__ samples,
                                                             const float* __restrict__ means,
                                                             const float* __restrict__ quats,
                                                             const float* __restrict__ scales,
                                                             float* __restrict__ means_out,
                                                             float* __restrict__ scales_out) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * RB + threadIdx.x;
    if (i >= n_children) return;
    const long long d = n_keep + i;
    const long long s = src[d], r = sample_row[d];
    if (r == -1) return;
    if (r == -2) {
        for (int j = 0; j < 3; ++j) scales_out[3 * d + j] = logf(expf(scales[3 * s + j]) / 1.6f);
        return;
    }
    float q[4] = {quats[4 * s], quats[4 * s + 1], quats[4 * s + 2], quats[4 * s + 3]};
    // the reference normalises, then quat_to_rotmat normalises again (:523-524)
    for (int rep = 0; rep < 2; ++rep) {
        const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        const float dnm = rep == 0 ? nrm : fmaxf(nrm, 1e-12f);  // F.normalize clamps its denominator
        for (int j = 0; j < 4; ++j) q[j] = __fdiv_rn(q[j], dnm);
    }
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float R[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - w * z), 2.f * (x * z + w * y),
                        2.f * (x * y + w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - w * x),
                        2.f * (x * z - w * y), 2.f * (y * z + w * x), 1.f - 2.f * (x * x + y * y)};
    float v[3], es[3];
    for (int j = 0; j < 3; ++j) {
        es[j] = expf(scales[3 * s + j]);
        v[j] = es[j] * samples[3 * r + j];
    }
    for (int j = 0; j < 3; ++j) {
        means_out[3 * d + j] = (R[3 * j] * v[0] + R[3 * j + 1] * v[1] + R[3 * j + 2] * v[2]) + means[3 * s + j];
        scales_out[3 * d + j] = logf(es[j] / 1.6f);
    }
}

}  // namespace
}  // namespace fg

using namespace fg;

extern "C" int64_t fg_refine_workspace_bytes(int64_t N) {
    return 4 * N * (int64_t)sizeof(int) + 256 + fg_scan_workspace_bytes(4 * N) + 64;
}

extern "C" int fg_refine_plan(int64_t N, const float* scales, const float* opacities, const float* grad_norm,
                              const float* vis_count, const float* max_size, const fg_refine_config* cfg,
                              int32_t* plan, int64_t* counts, void* workspace, int64_t workspace_bytes,
                              void* stream) {
    FG_REQUIRE(N >= 0 && 4 * N + 1 < (1ll << 31), "N out of range");
    FG_REQUIRE(cfg && plan && counts, "NULL pointer");
    FG_REQUIRE(cfg->n_split_samples >= 1, "n_split_samples");
    if (workspace_bytes < fg_refine_workspace_bytes(N))
        return set_error(FG_ERR_WORKSPACE, "workspace too small", __FILE__, __LINE__);
    FG_REQUIRE(workspace, "NULL pointer");
    if (N > 0) {
        FG_REQUIRE(scales && opacities, "NULL pointer");
        FG_REQUIRE(!cfg->densify || (grad_norm && vis_count), "densify needs the statistics");
    }
    char* w = (char*)workspace;
    int* flags = (int*)w;
    w += (4 * N * sizeof(int) + 255) / 256 * 256;
    long long* total = (long long*)w;
    w += 64;
    PlanParams p{N, scales, opacities, grad_norm, vis_count, max_size, *cfg, flags};
    if (N > 0) FG_LAUNCH(refine_flags_kernel, ceil_div(N, RB), RB, 0, stream, p);
    int e = fg_exclusive_scan_i32(4 * N, flags, plan, (int64_t*)total, w, workspace_bytes - (w - (char*)workspace),
                                  stream);
    if (e) return e;
    FG_LAUNCH(refine_counts_kernel, 1, 1, 0, stream, (long long)N, plan, total, (long long*)counts);
    return FG_OK;
}

extern "C" int fg_refine_map(int64_t N, const int32_t* plan, const int64_t* counts, int n_split_samples,
                             int64_t n_out, int32_t* src, int32_t* sample_row, void* stream) {
    FG_REQUIRE(N >= 0 && n_out >= 0 && n_split_samples >= 1, "shape");
    if (N == 0 || n_out == 0) return FG_OK;
    FG_REQUIRE(plan && counts && src && sample_row, "NULL pointer");
    FG_LAUNCH(refine_map_kernel, ceil_div(N, RB), RB, 0, stream, (long long)N, plan, (const long long*)counts,
              n_split_samples, (long long)n_out, src, sample_row);
    return FG_OK;
}

extern "C" int fg_refine_gather(int64_t n_out, int64_t n_keep, const int32_t* src, int n_arrays,
                                const fg_refine_array* arrays, void* stream) {
    FG_REQUIRE(n_arrays >= 0 && n_arrays <= FG_REFINE_MAX_ARRAYS, "array count");
    FG_REQUIRE(n_out >= 0 && n_keep >= 0 && n_keep <= n_out, "shape");
    if (n_out == 0 || n_arrays == 0) return FG_OK;
    FG_REQUIRE(src && arrays, "NULL pointer");
    GatherParams p{};
    p.n_out = n_out; p.n_keep = n_keep; p.src = src;
    int widest = 1;
    for (int i = 0; i < n_arrays; ++i) {
        FG_REQUIRE(arrays[i].in && arrays[i].out && arrays[i].row_floats >= 1, "array");
        p.a[i] = arrays[i];
        widest = arrays[i].row_floats > widest ? arrays[i].row_floats : widest;
    }
    long long blocks = (n_out * widest + RB - 1) / RB;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)n_arrays);
    FG_LAUNCH(refine_gather_kernel, grid, RB, 0, stream, p);
    return FG_OK;
}

extern "C" int fg_refine_children(int64_t n_out, int64_t n_keep, int64_t n_children, const int32_t* src,
                                  const int32_t* sample_row, const float* samples, const float* means,
                                  const float* quats, const float* scales, float* means_out, float* scales_out,
                                  void* stream) {
    FG_REQUIRE(n_keep >= 0 && n_children >= 0 && n_keep + n_children <= n_out, "shape");
    const long long n_rows = n_out - n_keep;  // duplicates of split parents take the rescaled scale too
    if (n_rows == 0) return FG_OK;
    FG_REQUIRE(src && sample_row && means && quats && scales && means_out && scales_out, "NULL pointer");
    FG_REQUIRE(n_children == 0 || samples, "NULL pointer");
    FG_LAUNCH(refine_children_kernel, ceil_div(n_rows, RB), RB, 0, stream, (long long)n_keep, n_rows, src, sample_row, samples, means, quats, scales, means_out, scales_out);
    return FG_OK;
}
