// (4) Exact k-nearest neighbours of a 3-D point set against itself on a uniform grid.
// Replaces FreeGaussianModel.k_nearest_sklearn (freegaussian_model.py:293-311; sklearn
// NearestNeighbors(n_neighbors=k+1, metric="euclidean"), column 0 dropped).
//
// Bit-exactness contract: distances are computed exactly as sklearn's KDTree does for
// float32 input -- promote to float64, d2 = (dx*dx + dy*dy) + dz*dz with each operation
// rounded (no FMA contraction), result sqrt(d2) rounded to float32 -- and neighbours are
// the k+1 smallest under the total order (d2, index); the first (self, for duplicate-free
// input) is dropped.  The grid only prunes: a ring of cells is skipped only when a
// conservative lower bound on its distance already exceeds the current k-th best, so the
// result equals exhaustive search.
//
// Roofline: HBM/L2 candidate streaming, 16 B per candidate point; FP64 pipe 8 flop per
// candidate.  Brute-force floor 8 N^2 flop.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

extern "C" int fg_exclusive_scan_i32(int64_t n, const int32_t* counts, int32_t* offsets, int64_t* total,
                                     void* workspace, int64_t workspace_bytes, void* stream);
extern "C" int64_t fg_scan_workspace_bytes(int64_t n);

namespace fg {

struct KnnGrid {
    double ox, oy, oz;  // origin (bbox min)
    double inv_h, h;
    int nx, ny, nz;
};

__device__ __forceinline__ unsigned f2ord(float f) {  // order-preserving float -> uint
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u) {
    unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

__global__ void knn_bbox_kernel(long long n, const float* __restrict__ pts, unsigned* __restrict__ bbox /*[6]*/) {
    pdl_wait();
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            unsigned o = f2ord(pts[3 * i + a]);
            lo[a] = min(lo[a], o);
            hi[a] = max(hi[a], o);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(bbox + a, lo[a]);
            atomicMax(bbox + 3 + a, hi[a]);
        }
    }
}

__device__ __forceinline__ int cell_coord(double p, double o, double inv_h, int n) {
    int c = (int)floor((p - o) * inv_h);
    return max(0, min(n - 1, c));
}

__global__ void knn_count_kernel(long long n, const float* __restrict__ pts, KnnGrid g, int32_t* __restrict__ cell_count,
                                 int32_t* __restrict__ cell_of) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_coord((double)pts[3 * i], g.ox, g.inv_h, g.nx);
    int cy = cell_coord((double)pts[3 * i + 1], g.oy, g.inv_h, g.ny);
    int cz = cell_coord((double)pts[3 * i + 2], g.oz, g.inv_h, g.nz);
    int cid = (cz * g.ny + cy) * g.nx + cx;
    cell_of[i] = cid;
    atomicAdd(cell_count + cid, 1);
}

__global__ void knn_scatter_kernel(long long n, const float* __restrict__ pts, const int32_t* __restrict__ cell_of,
                                   const int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_fill,
                                   float4* __restrict__ sorted) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cid = cell_of[i];
    int pos = cell_start[cid] + atomicAdd(cell_fill + cid, 1);
    sorted[pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float((int)i));
}

// Sorted list of the KCAP best (d2, idx), ascending; for KCAP >= 17 held in REGISTERS: every index is a compile-time
// constant after unrolling.  (The first version indexed the arrays with run-time slots, which put them in local memory: ncu r2
// showed 6.1 GB of DRAM writes and 30 long-scoreboard stalls per issue for a kernel whose output is 0.38 GB.)  The list
// always has KCAP slots; a query for fewer neighbours reads a prefix.
template <int KCAP>
struct BestList {
    double d2[KCAP];
    int idx[KCAP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KCAP; ++i) { d2[i] = INFINITY; idx[i] = 0x7fffffff; }
    }
    __device__ __forceinline__ bool worse_than(double d, int i, double dj, int ij) const {
        return d < dj || (d == dj && i < ij);  // (d, i) precedes (dj, ij) in the total order (d2, index)
    }
    __device__ __forceinline__ void insert(double d, int i) {
        if (!worse_than(d, i, d2[KCAP - 1], idx[KCAP - 1])) return;
        if (KCAP <= 9) {
            // short lists: a plain shifting loop (run-time slots, the few entries sit in L1-resident local memory and
            // the kernel keeps 16 blocks per SM); measured 1.4 ms against 1.8-2.3 ms for the unrolled form at k = 3
            int s = KCAP - 1;
            while (s > 0 && worse_than(d, i, d2[s - 1], idx[s - 1])) {
                d2[s] = d2[s - 1];
                idx[s] = idx[s - 1];
                --s;
            }
            d2[s] = d;
            idx[s] = i;
            return;
        }
        // one pass from the back: slot s takes its left neighbour while the new element precedes that neighbour,
        // the new element lands in the first slot whose left neighbour it does not precede
        bool placed = false;
#pragma unroll
        for (int s = KCAP - 1; s > 0; --s) {
            const bool before = worse_than(d, i, d2[s - 1], idx[s - 1]);
            if (!placed) {
                if (before) { d2[s] = d2[s - 1]; idx[s] = idx[s - 1]; }
                else { d2[s] = d; idx[s] = i; placed = true; }
            }
        }
        if (!placed) { d2[0] = d; idx[0] = i; }
    }
    __device__ __forceinline__ double dist2_at(int k) const {  // run-time slot without dynamic indexing
        double v = d2[0];
#pragma unroll
        for (int s = 1; s < KCAP; ++s) v = (s == k) ? d2[s] : v;
        return v;
    }
};

template <int KCAP>
__global__ void __launch_bounds__(128, (KCAP <= 9 ? 8 : KCAP <= 17 ? 4 : 1))
    knn_query_kernel(long long n, const float4* __restrict__ sorted, const int32_t* __restrict__ cell_start,
                     KnnGrid g, int k, float* __restrict__ out_dist, int32_t* __restrict__ out_idx) {
    pdl_wait();
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float4 me = sorted[q];
    const double qx = me.x, qy = me.y, qz = me.z;
    const int self = __float_as_int(me.w);
    const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx);
    const int cy = cell_coord(qy, g.oy, g.inv_h, g.ny);
    const int cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
    BestList<KCAP> best;
    best.init();
    const int rmax = max(max(g.nx, g.ny), g.nz);
    for (int r = 0; r <= rmax; ++r) {
        // cells at Chebyshev distance exactly r from (cx,cy,cz)
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
        for (int z = z0; z <= z1; ++z) {
            const bool zface = (z == cz - r) || (z == cz + r);
            for (int y = y0; y <= y1; ++y) {
                const bool yface = (y == cy - r) || (y == cy + r);
                // a face row of the shell is walked in full, an interior row only at its two x faces; ONE loop body, so
                // that the candidate scan is inlined once and the best list stays in registers
                int xa = x0, xb = x1, xstep = 1;
                if (!(zface || yface)) { xa = cx - r; xb = cx + r; xstep = 2 * r; }  // r >= 1 here (r == 0 is all faces)
                for (int x = xa; x <= xb; x += xstep) {
                    if (x < 0 || x >= g.nx) continue;
                    const int cid = (z * g.ny + y) * g.nx + x;
                    const int s = cell_start[cid], e = cell_start[cid + 1];
                    for (int j = s; j < e; ++j) {
                        const float4 c = sorted[j];
                        const double dx = __dsub_rn(qx, (double)c.x);
                        const double dy = __dsub_rn(qy, (double)c.y);
                        const double dz = __dsub_rn(qz, (double)c.z);
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        best.insert(d2, __float_as_int(c.w));
                    }
                }
            }
        }
        // every unvisited point differs by more than r cells along some axis, hence lies
        // farther than r*h (up to rounding, absorbed by the 1e-9 slack)
        const double lb = (double)r * g.h * (1.0 - 1e-9);
        if (best.dist2_at(k) <= lb * lb) break;
        if (x0 == 0 && y0 == 0 && z0 == 0 && x1 == g.nx - 1 && y1 == g.ny - 1 && z1 == g.nz - 1) break;  // whole grid seen
    }
    // drop column 0 (self for duplicate-free input), as freegaussian_model.py:311 does
#pragma unroll
    for (int j = 0; j < KCAP - 1; ++j) {
        if (j < k) {
            out_dist[(size_t)self * k + j] = (float)sqrt(best.d2[j + 1]);
            out_idx[(size_t)self * k + j] = best.idx[j + 1];
        }
    }
}

struct KnnLayout {
    size_t off_bbox, off_cell_of, off_sorted, off_cells, off_scan, total;
    long long max_cells;
};
static KnnLayout knn_layout(long long n) {
    KnnLayout L;
    L.max_cells = 2 * n + 1024;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) & ~(size_t)255; return r; };
    L.off_bbox = take(6 * 4 + 8);
    L.off_cell_of = take((size_t)n * 4);
    L.off_sorted = take((size_t)n * 16);
    L.off_cells = take((size_t)(L.max_cells + 1) * 4 * 3);  // count, start, fill
    L.off_scan = take((size_t)fg_scan_workspace_bytes(L.max_cells + 1));
    L.total = o;
    return L;
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_knn_workspace_bytes(int64_t n) { return (int64_t)knn_layout(n < 1 ? 1 : n).total; }

extern "C" int fg_knn_f32(int64_t n, const float* points, int k, float* out_dist, int32_t* out_idx, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    FG_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
    FG_REQUIRE(k >= 1 && k <= 64, "k must be in 1..64");
    FG_REQUIRE(n == 0 || n > k, "need more than k points");
    if (n == 0) return FG_OK;
    FG_REQUIRE(points && out_dist && out_idx && workspace, "NULL pointer");
    KnnLayout L = knn_layout(n);
    FG_REQUIRE((size_t)workspace_bytes >= L.total, "knn workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    unsigned* bbox = (unsigned*)(ws + L.off_bbox);
    int32_t* cell_of = (int32_t*)(ws + L.off_cell_of);
    float4* sorted = (float4*)(ws + L.off_sorted);
    int32_t* cell_count = (int32_t*)(ws + L.off_cells);
    int32_t* cell_start = cell_count + (L.max_cells + 1);
    int32_t* cell_fill = cell_start + (L.max_cells + 1);
    int64_t* scan_total = (int64_t*)(ws + L.off_bbox + 24);

    // 1. bounding box (host needs it to size the grid: one small D2H + sync, init-time only)
    unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    FG_CUDA(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
    FG_LAUNCH(knn_bbox_kernel, (int)std::min<long long>((long long)num_sms() * 8, ((long long)n + 255) / 256), 256, 0, st, (long long)n, points, bbox);
    unsigned hb[6];
    FG_CUDA(cudaMemcpyAsync(hb, bbox, sizeof(hb), cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaStreamSynchronize(st));
    double lo[3], ext[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = (double)ord2f(hb[a]);
        ext[a] = (double)ord2f(hb[3 + a]) - lo[a];
        FG_REQUIRE(isfinite(lo[a]) && isfinite(ext[a]), "points contain non-finite values");
    }
    // 2. cell size: ~max(2, (k+1)/2) points per cell in the occupied volume; never more than 2n+1024 cells
    const double ppc = fmax(2.0, 0.5 * (k + 1));
    double dims_used = 0, vol = 1.0;
    for (int a = 0; a < 3; ++a)
        if (ext[a] > 0) { vol *= ext[a]; dims_used += 1; }
    double h = dims_used > 0 ? pow(vol * ppc / (double)n, 1.0 / dims_used) : 1.0;
    if (!(h > 0) || !isfinite(h)) h = 1.0;
    KnnGrid g;
    while (true) {
        double nx = floor(ext[0] / h) + 1, ny = floor(ext[1] / h) + 1, nz = floor(ext[2] / h) + 1;
        if (nx * ny * nz <= (double)L.max_cells && nx < 2e9 && ny < 2e9 && nz < 2e9) {
            g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz;
            break;
        }
        h *= 1.2599210498948732;
    }
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.h = h; g.inv_h = 1.0 / h;
    const long long cells = (long long)g.nx * g.ny * g.nz;
    // 3. counting sort of the points by cell
    FG_CUDA(cudaMemsetAsync(cell_count, 0, (size_t)(L.max_cells + 1) * 4 * 3, st));
    const int nblk = ceil_div(n, 256);
    FG_LAUNCH(knn_count_kernel, nblk, 256, 0, st, (long long)n, points, g, cell_count, cell_of);
    if (int e = fg_exclusive_scan_i32(cells + 1, cell_count, cell_start, scan_total, ws + L.off_scan,
                                      (int64_t)(L.total - L.off_scan), stream))
        return e;
    FG_LAUNCH(knn_scatter_kernel, nblk, 256, 0, st, (long long)n, points, cell_of, cell_start, cell_fill, sorted);
    // 4. query
    const int qblk = ceil_div(n, 128);
    if (k + 1 <= 4) {
        FG_LAUNCH((knn_query_kernel<4>), qblk, 128, 0, st, (long long)n, sorted, cell_start, g, k, out_dist, out_idx);
    } else if (k + 1 <= 9) {
        FG_LAUNCH((knn_query_kernel<9>), qblk, 128, 0, st, (long long)n, sorted, cell_start, g, k, out_dist, out_idx);
    } else if (k + 1 <= 17) {
        FG_LAUNCH((knn_query_kernel<17>), qblk, 128, 0, st, (long long)n, sorted, cell_start, g, k, out_dist, out_idx);
    } else if (k + 1 <= 33) {
        FG_LAUNCH((knn_query_kernel<33>), qblk, 128, 0, st, (long long)n, sorted, cell_start, g, k, out_dist, out_idx);
    } else {
        FG_LAUNCH((knn_query_kernel<65>), qblk, 128, 0, st, (long long)n, sorted, cell_start, g, k, out_dist, out_idx);
    }
    return FG_OK;
}
