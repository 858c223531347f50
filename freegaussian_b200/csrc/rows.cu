// Host-sync-free helpers of the deformation network's sparse backward and its time branch.
//
// (a) Gaussians that were culled or never reached a pixel in this step's views get an exactly-zero gradient, so the
//     network backward (freegaussian_model.py:1054-1114 through autograd) only needs the other rows.  fg_rows_active
//     compacts their indices on the device (flags -> scan -> scatter, ascending order) and leaves the COUNT on the
//     device; fg_rows_gather copies the selected rows of a saved activation into a buffer of a capacity the host guessed
//     from the previous step, zero-filling the rows past the count -- zero rows add exactly nothing to any product of the
//     backward, so no host read is needed before the tensor-core kernels are enqueued.
// (b) The time branch (positional embedding of ONE time value + `timenet`, freegaussian_model.py:1066-1071, 1094-1096):
//     one 256-thread block forward and one backward instead of ~60 single-row torch launches per iteration.
#include <algorithm>

#include "common.cuh"

extern "C" int fg_exclusive_scan_i32(int64_t n, const int32_t* counts, int32_t* offsets, int64_t* total,
                                     void* workspace, int64_t workspace_bytes, void* stream);
extern "C" int64_t fg_scan_workspace_bytes(int64_t n);

namespace fg {

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// one warp per row: flag = any element != 0
__global__ void __launch_bounds__(256) rows_flag_kernel(long long N, const float* __restrict__ g, int ld, int32_t* __restrict__ flags) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (row >= N) return;
    const float* r = g + row * ld;
    bool nz = false;
    for (int c = lane; c < ld; c += 32) nz |= r[c] != 0.f;
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) flags[row] = nz ? 1 : 0;
}

__global__ void __launch_bounds__(256) rows_scatter_kernel(long long N, const int32_t* __restrict__ flags,
                                                           const int32_t* __restrict__ offsets, int32_t* __restrict__ idx) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < N && flags[i]) idx[offsets[i]] = (int32_t)i;
}

// rows of `units` 16-byte words; thread = one word
__global__ void __launch_bounds__(256) rows_gather_kernel(long long M, const int32_t* __restrict__ idx, const long long* __restrict__ count,
                                                          const uint4* __restrict__ src, int units, uint4* __restrict__ dst) {
    pdl_wait();
    const long long n = *count;
    const long long total = M * units;
    for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < total; q += (long long)gridDim.x * 256) {
        const long long r = q / units;
        const int u = (int)(q - r * units);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < n) v = src[(long long)idx[r] * units + u];
        dst[q] = v;
    }
}

// ---- time branch: emb = [t, sin(2^k t), cos(2^k t)]_k ; h = relu(W1 emb + b1) ; out = W2 h + b2 (or out = emb) -------
__global__ void __launch_bounds__(256) time_branch_fwd_kernel(const float* __restrict__ t, int multires, int in_ch, int hidden,
                                                              int out_ch, const float* __restrict__ w1, const float* __restrict__ b1,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              float* __restrict__ emb, float* __restrict__ h, float* __restrict__ out) {
    pdl_wait();
    __shared__ float se[64];
    __shared__ float sh[256];
    const int tid = threadIdx.x;
    const float tv = t[0];
    if (tid < in_ch) {
        float v = tv;
        if (tid > 0) {
            const int k = (tid - 1) >> 1;
            const float a = tv * exp2f((float)k);
            v = ((tid - 1) & 1) ? cosf(a) : sinf(a);
        }
        se[tid] = v;
        emb[tid] = v;
    }
    __syncthreads();
    if (w1 == nullptr) return;  // no timenet: the embedding is the branch's output
    if (tid < hidden) {
        float a = b1[tid];
        for (int i = 0; i < in_ch; ++i) a = fmaf(w1[tid * in_ch + i], se[i], a);
        a = fmaxf(a, 0.f);
        sh[tid] = a;
        h[tid] = a;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int o = warp; o < out_ch; o += 8) {
        float a = 0.f;
        for (int j = lane; j < hidden; j += 32) a = fmaf(w2[o * hidden + j], sh[j], a);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) out[o] = a + b2[o];
    }
}

__global__ void __launch_bounds__(256) time_branch_bwd_kernel(int in_ch, int hidden, int out_ch, const float* __restrict__ emb,
                                                              const float* __restrict__ h, const float* __restrict__ w2,
                                                              const float* __restrict__ g_out, float* __restrict__ dw1,
                                                              float* __restrict__ db1, float* __restrict__ dw2,
                                                              float* __restrict__ db2) {
    pdl_wait();
    __shared__ float sg[64];
    __shared__ float se[64];
    const int tid = threadIdx.x;
    if (tid < out_ch) { sg[tid] = g_out[tid]; db2[tid] = g_out[tid]; }
    if (tid < in_ch) se[tid] = emb[tid];
    __syncthreads();
    if (tid < hidden) {
        const float hv = h[tid];
        float dh = 0.f;
        for (int o = 0; o < out_ch; ++o) {
            dw2[o * hidden + tid] = sg[o] * hv;
            dh = fmaf(w2[o * hidden + tid], sg[o], dh);
        }
        if (!(hv > 0.f)) dh = 0.f;
        db1[tid] = dh;
        for (int i = 0; i < in_ch; ++i) dw1[tid * in_ch + i] = dh * se[i];
    }
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_rows_workspace_bytes(int64_t N) {
    const size_t n = (size_t)(N < 1 ? 1 : N);
    return (int64_t)(al256(n * 4) * 2 + al256((size_t)fg_scan_workspace_bytes((int64_t)n)));
}

extern "C" int fg_rows_active(int64_t N, const float* g, int ld, int32_t* idx, int64_t* count_dev, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    FG_REQUIRE(N >= 0 && N < (1ll << 31) && ld >= 1, "bad N / ld");
    FG_REQUIRE(count_dev != nullptr, "count_dev must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        FG_CUDA(cudaMemsetAsync(count_dev, 0, 8, st));
        return FG_OK;
    }
    FG_REQUIRE(g && idx && workspace, "NULL pointer");
    FG_REQUIRE(workspace_bytes >= fg_rows_workspace_bytes(N), "rows workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    int32_t* flags = (int32_t*)ws;
    int32_t* offsets = (int32_t*)(ws + al256((size_t)N * 4));
    unsigned char* scan_ws = ws + 2 * al256((size_t)N * 4);
    FG_CUDA(cudaMemsetAsync(idx, 0, (size_t)N * 4, st));
    FG_LAUNCH(rows_flag_kernel, ceil_div(N * 32, 256), 256, 0, st, (long long)N, g, ld, flags);
    if (int e = fg_exclusive_scan_i32(N, flags, offsets, count_dev, scan_ws, fg_scan_workspace_bytes(N), stream)) return e;
    FG_LAUNCH(rows_scatter_kernel, ceil_div(N, 256), 256, 0, st, (long long)N, flags, offsets, idx);
    return FG_OK;
}

extern "C" int fg_rows_gather(int64_t M, const int32_t* idx, const int64_t* count_dev, const void* src, int row_bytes,
                              void* dst, void* stream) {
    FG_REQUIRE(M >= 0 && row_bytes > 0 && row_bytes % 16 == 0, "row_bytes must be a positive multiple of 16");
    if (M == 0) return FG_OK;
    FG_REQUIRE(idx && count_dev && src && dst, "NULL pointer");
    FG_REQUIRE(((uintptr_t)src | (uintptr_t)dst) % 16 == 0, "src / dst must be 16-byte aligned");
    const int units = row_bytes / 16;
    const int grid = (int)std::min<long long>(((long long)M * units + 255) / 256, (long long)num_sms() * 16);
    FG_LAUNCH(rows_gather_kernel, grid, 256, 0, (cudaStream_t)stream, (long long)M, idx, (const long long*)count_dev,
              (const uint4*)src, units, (uint4*)dst);
    return FG_OK;
}

extern "C" int fg_time_branch_fwd(const float* t, int multires, int in_ch, int hidden, int out_ch, const float* w1,
                                  const float* b1, const float* w2, const float* b2, float* emb, float* h, float* out,
                                  void* stream) {
    FG_REQUIRE(t && emb, "NULL pointer");
    FG_REQUIRE(in_ch == 1 + 2 * multires && in_ch <= 64, "in_ch must be 1 + 2 * multires and <= 64");
    if (w1) {
        FG_REQUIRE(hidden >= 1 && hidden <= 256 && out_ch >= 1 && out_ch <= 64, "hidden must be <= 256 and out_ch <= 64");
        FG_REQUIRE(b1 && w2 && b2 && h && out, "NULL pointer");
    }
    FG_LAUNCH(time_branch_fwd_kernel, 1, 256, 0, (cudaStream_t)stream, t, multires, in_ch, hidden, out_ch, w1, b1, w2, b2, emb, h, out);
    return FG_OK;
}

extern "C" int fg_time_branch_bwd(int in_ch, int hidden, int out_ch, const float* emb, const float* h, const float* w2,
                                  const float* g_out, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
    FG_REQUIRE(in_ch >= 1 && in_ch <= 64 && hidden >= 1 && hidden <= 256 && out_ch >= 1 && out_ch <= 64, "bad sizes");
    FG_REQUIRE(emb && h && w2 && g_out && dw1 && db1 && dw2 && db2, "NULL pointer");
    FG_LAUNCH(time_branch_bwd_kernel, 1, 256, 0, (cudaStream_t)stream, in_ch, hidden, out_ch, emb, h, w2, g_out, dw1, db1, dw2, db2);
    return FG_OK;
}
