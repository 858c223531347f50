// (2c) Hierarchical tile binning: builds the per-tile depth-sorted lists WITHOUT sorting the
// tile intersections.  Same output (flatten_ids, isect_offsets) as the reference's 64-bit key
// sort: inside a tile, entries are ordered by (depth bits, c*N+n).
//
//   1. the (c,n) splats are stably sorted by depth once (radix_sort.cu, 32-bit keys)   [N items]
//   2. bin_count: per splat, +1/-1 at the four corners of its tile rectangle in a 2-D
//      difference grid, and the number of COARSE cells (4x4 tiles) it overlaps
//   3. tile_scan: 2-D prefix sum of the grid = exact intersections per tile; their exclusive
//      scan = isect_offsets; the grand total M sizes flatten_ids
//   4. the (splat, coarse cell) pairs are emitted in depth order and stably sorted by cell
//      (isect.cu emit + radix sort on ~0.1 M items: an order of magnitude fewer than M)
//   5. fine_bin: one CTA per coarse cell streams the cell's list in order and appends each
//      splat to the lists of the (up to 16) tiles it overlaps with ballot-ranked, coalesced
//      stores.  Order inside a tile = order of the stream = depth order: stable by construction.
//
// Roofline: HBM.  Algorithmic bytes: 20 B per splat (count) + 28 B per coarse pair + 4 B per
// tile intersection written once -- against 36 B (two-level sort) or 152 B (64-bit sort) per
// intersection.
#include "common.cuh"
#include "splat_math.h"

namespace fg {

constexpr int CK = 4;        // coarse cell = CK x CK tiles
constexpr int CK_SHIFT = 2;

// ---- step 2 -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    bin_count_kernel(int C, int N, long long total, const int32_t* __restrict__ order,
                     const float2* __restrict__ means2d, const int32_t* __restrict__ radii, int tile_size, int tile_w,
                     int tile_h, int32_t* __restrict__ diff /*[C][tile_h+1][tile_w+1]*/,
                     int32_t* __restrict__ coarse_cnt /*[total], in `order`*/) {
    pdl_wait();
    const long long slot = (long long)blockIdx.x * 256 + threadIdx.x;
    if (slot >= total) return;
    const long long idx = order[slot];
    if (idx < 0) {  // past the visible splats (fg_depth_sort_visible leaves -1 there)
        coarse_cnt[slot] = 0;
        return;
    }
    const int r = radii[idx];
    int cnt = 0;
    if (r > 0) {
        const float2 m = means2d[idx];
        const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h);
        if (t.x1 > t.x0 && t.y1 > t.y0) {
            int32_t* g = diff + (idx / N) * (long long)(tile_h + 1) * (tile_w + 1);
            const int W1 = tile_w + 1;
            atomicAdd(g + t.y0 * W1 + t.x0, 1);
            atomicAdd(g + t.y0 * W1 + t.x1, -1);
            atomicAdd(g + t.y1 * W1 + t.x0, -1);
            atomicAdd(g + t.y1 * W1 + t.x1, 1);
            cnt = (((t.x1 + CK - 1) >> CK_SHIFT) - (t.x0 >> CK_SHIFT)) * (((t.y1 + CK - 1) >> CK_SHIFT) - (t.y0 >> CK_SHIFT));
        }
    }
    coarse_cnt[slot] = cnt;
}

// ---- step 3: one CTA per camera turns its difference grid into per-tile counts (2-D prefix sum
// in place, then compacted to [tile_h][tile_w]); the exclusive scan over all tiles is the generic
// scan of isect.cu, so many-camera calls (cfg4: 32 views, ~0.7 M tiles) stay parallel.
__global__ void __launch_bounds__(1024)
    tile_count_kernel(int tile_w, int tile_h, int32_t* __restrict__ diff, int32_t* __restrict__ counts,
                      int32_t* __restrict__ offsets_out /*one camera only: the exclusive scan too, or NULL*/,
                      int64_t* __restrict__ total_out) {
    __shared__ int s_scan[32];
    pdl_wait();
    const int W1 = tile_w + 1, H1 = tile_h + 1;
    const int c = blockIdx.x;
    int32_t* g = diff + (long long)c * H1 * W1;
    const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // row-wise prefix (warp per row, coalesced)
    for (int y = warp_; y < H1; y += nwarps) {
        int carry = 0;
        for (int x0 = 0; x0 < W1; x0 += 32) {
            const int x = x0 + lane_;
            int v = x < W1 ? g[y * W1 + x] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane_ >= d) v += o;
            }
            v += carry;
            if (x < W1) g[y * W1 + x] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // column-wise prefix (thread per column, next row prefetched) + compaction
    int32_t* out = counts + (long long)c * tile_h * tile_w;
    for (int x = threadIdx.x; x < tile_w; x += blockDim.x) {
        int run = 0;
        int next = g[x];
        for (int y = 0; y < tile_h; ++y) {
            const int cur = next;
            next = g[(y + 1) * W1 + x];
            run += cur;
            out[y * tile_w + x] = run;
        }
    }
    if (!offsets_out) return;
    // single camera: this CTA holds every count, so it finishes the job (three launches of the generic scan otherwise)
    __syncthreads();
    const int n = tile_w * tile_h, per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int i0 = min((int)threadIdx.x * per, n), i1 = min(i0 + per, n);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += out[i];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane_ >= d) incl += o;
    }
    if (lane_ == 31) s_scan[warp_] = incl;
    __syncthreads();
    if (warp_ == 0) {
        int w = lane_ < nwarps ? s_scan[lane_] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane_ >= d) w += o;
        }
        s_scan[lane_] = w;  // inclusive over warps
    }
    __syncthreads();
    int run = incl - sum + (warp_ ? s_scan[warp_ - 1] : 0);
    for (int i = i0; i < i1; ++i) {
        const int c = out[i];
        offsets_out[i] = run;
        run += c;
    }
    if (threadIdx.x == blockDim.x - 1) *total_out = run;
}

// ---- step 4: emit (splat, coarse cell) pairs in depth order --------------------------------
__global__ void __launch_bounds__(256)
    coarse_emit_kernel(int C, int N, long long total, const int32_t* __restrict__ order,
                       const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                       const int32_t* __restrict__ coarse_off, int tile_size, int tile_w, int tile_h, int cw, int chh,
                       uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    pdl_wait();
    const long long slot = (long long)blockIdx.x * 256 + threadIdx.x;
    if (slot >= total) return;
    const long long idx = order[slot];
    if (idx < 0) return;  // past the visible splats
    const int r = radii[idx];
    if (r <= 0) return;
    const float2 m = means2d[idx];
    const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h);
    if (!(t.x1 > t.x0 && t.y1 > t.y0)) return;
    const int cx0 = t.x0 >> CK_SHIFT, cx1 = (t.x1 + CK - 1) >> CK_SHIFT;
    const int cy0 = t.y0 >> CK_SHIFT, cy1 = (t.y1 + CK - 1) >> CK_SHIFT;
    const uint32_t cam_base = (uint32_t)((idx / N) * (long long)cw * chh);
    int k = coarse_off[slot];
    for (int y = cy0; y < cy1; ++y)
        for (int x = cx0; x < cx1; ++x) {
            keys[k] = cam_base + (uint32_t)(y * cw + x);
            vals[k] = (int32_t)idx;
            ++k;
        }
}

// ---- steps 2 + 4 without a sort: ranked placement of the (splat, cell) pairs -----------------------------------------
// The pairs only have to end up grouped by cell, in depth order inside a cell.  With at most RK_MAX_CELLS coarse cells
// (one 1080p view: 510) that position can be COMPUTED instead of sorted for:
//   bin_count_cells : CTA = chunk of RK_CHUNK consecutive depth-ordered slots; besides the corner increments it counts
//                     the chunk's pairs per cell (a difference grid in cell space) -> mat[chunk][cell]
//   cell_scan       : per cell, exclusive prefix over the chunks (in place) and the cell totals; the last CTA to finish
//                     scans the totals -> cell_offsets[n_cells + 1], Mc
//   ranked_emit     : CTA = chunk again; a bitmap per (cell, warp) of the chunk's splats that touch the cell; position of
//                     a pair = cell_offsets[cell] + mat[chunk][cell] + (set bits below this splat): depth order, exactly
//                     what the stable sort by cell produced -- without keys, histogram, two onesweep passes, the scan over
//                     C*N per-splat counts and the cell-offsets kernel (7 launches, 0.12 ms of 1.56 at cfg3).
constexpr int RK_CHUNK = 512;
constexpr int RK_WARPS = RK_CHUNK / 32;
constexpr int RK_MAX_CELLS = 1024;  // = threads of the totals scan; shared memory of ranked_emit: 128 B per cell
constexpr int RK_BIG = 16;  // a splat over more cells than this is walked by its whole warp (a full-screen splat touches every
                           // cell: one thread looping over 510 of them would hold up the CTA's barriers)

// cells [x0, x0 + w) x [y0, ..) of a cell rectangle with n cells, as the warp sees the rectangle of lane `src`
struct WarpRect {
    int w, n, base;     // width, cells, index of the rectangle's first cell
    unsigned inv_w;     // ceil(2^20 / w): i / w == (i * inv_w) >> 20 for i, w <= 1024 (i * w < 2^20)
    __device__ __forceinline__ int cell(int i, int cw) const {  // index of the rectangle's i-th cell (row-major)
        const int y = (int)(((unsigned)i * inv_w) >> 20);
        return base + y * cw + (i - y * w);
    }
};
__device__ __forceinline__ WarpRect warp_rect(int src, int w, int n, int base) {
    WarpRect r;
    r.w = __shfl_sync(0xffffffffu, w, src);
    r.n = __shfl_sync(0xffffffffu, n, src);
    r.base = __shfl_sync(0xffffffffu, base, src);
    r.inv_w = ((1u << 20) + (unsigned)r.w - 1) / (unsigned)r.w;
    return r;
}

struct RankedLayout {
    size_t mat, total, counter, bytes;
    long long n_chunks;
    int n_cells;
};
static RankedLayout ranked_layout(int C, int N, int tile_w, int tile_h) {
    RankedLayout L;
    const int cw = (tile_w + CK - 1) / CK, chh = (tile_h + CK - 1) / CK;
    const long long cells = (long long)C * cw * chh;
    L.n_cells = cells <= RK_MAX_CELLS ? (int)cells : 0;
    L.n_chunks = ((long long)C * N + RK_CHUNK - 1) / RK_CHUNK;
    size_t o = 0;
    L.mat = o; o += ((size_t)L.n_chunks * L.n_cells * 4 + 255) & ~(size_t)255;
    L.total = o; o += ((size_t)L.n_cells * 4 + 255) & ~(size_t)255;
    L.counter = o; o += 256;
    L.bytes = L.n_cells ? o : 0;
    return L;
}

// Pair counts per cell of one chunk.  A splat covers ~11 cells at cfg3, and one shared-memory atomic per (splat, cell) made
// this kernel MIO-bound (ncu r2b: 17 mio_throttle stall cycles per issue, 43 us).  The counts come from a difference grid
// in cell space instead -- four atomics per splat at the corners of its cell rectangle, then a 2-D prefix sum of the
// (cw+1) x (chh+1) grid of every camera -- the same trick as the tile counts, one level up.
__global__ void __launch_bounds__(RK_CHUNK)
    bin_count_cells_kernel(int C, int N, long long total, const int32_t* __restrict__ order,
                           const float2* __restrict__ means2d, const int32_t* __restrict__ radii, int tile_size, int tile_w,
                           int tile_h, int cw, int chh, int n_cells, int32_t* __restrict__ diff, int32_t* __restrict__ mat,
                           int32_t* __restrict__ counter) {
    extern __shared__ int s_grid[];  // [C][chh + 1][cw + 1]
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long slot0 = (long long)blockIdx.x * RK_CHUNK;
    if (blockIdx.x == 0 && tid == 0) *counter = 0;  // cell_scan's "last CTA" ticket
    if (order[slot0] < 0) return;  // the whole chunk lies past the visible splats (fg_depth_sort_visible leaves -1 there)
    const int GW = cw + 1, GH = chh + 1;
    for (int i = tid; i < C * GW * GH; i += RK_CHUNK) s_grid[i] = 0;
    __syncthreads();
    const long long slot = slot0 + tid;
    const long long idx = slot < total ? order[slot] : -1;
    if (idx >= 0) {
        const int r = radii[idx];
        if (r > 0) {
            const float2 m = means2d[idx];
            const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h);
            if (t.x1 > t.x0 && t.y1 > t.y0) {
                const int cam = (int)(idx / N);
                int32_t* g = diff + cam * (long long)(tile_h + 1) * (tile_w + 1);
                const int W1 = tile_w + 1;
                atomicAdd(g + t.y0 * W1 + t.x0, 1);
                atomicAdd(g + t.y0 * W1 + t.x1, -1);
                atomicAdd(g + t.y1 * W1 + t.x0, -1);
                atomicAdd(g + t.y1 * W1 + t.x1, 1);
                const int cx0 = t.x0 >> CK_SHIFT, cx1 = (t.x1 + CK - 1) >> CK_SHIFT;
                const int cy0 = t.y0 >> CK_SHIFT, cy1 = (t.y1 + CK - 1) >> CK_SHIFT;
                int* c = s_grid + cam * GW * GH;
                atomicAdd(c + cy0 * GW + cx0, 1);
                atomicAdd(c + cy0 * GW + cx1, -1);
                atomicAdd(c + cy1 * GW + cx0, -1);
                atomicAdd(c + cy1 * GW + cx1, 1);
            }
        }
    }
    __syncthreads();
    // rows: a warp per row, 32 columns at a time with a carry
    for (int row = warp; row < C * GH; row += RK_WARPS) {
        int* g = s_grid + row * GW;
        int carry = 0;
        for (int x0 = 0; x0 < GW; x0 += 32) {
            const int x = x0 + lane;
            int v = x < GW ? g[x] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += o;
            }
            v += carry;
            if (x < GW) g[x] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // columns: a thread per (camera, column); the running sum is the cell's count
    int32_t* out = mat + (long long)blockIdx.x * n_cells;
    for (int col = tid; col < C * cw; col += RK_CHUNK) {
        const int cam = col / cw, x = col - cam * cw;
        const int* g = s_grid + cam * GW * GH + x;
        int run = 0;
        for (int y = 0; y < chh; ++y) {
            run += g[y * GW];
            out[(cam * chh + y) * cw + x] = run;
        }
    }
}

// grid = ceil(n_cells / 32) CTAs of 32 x 32 threads: lane = cell, row = a 1/32 share of the active chunks
__global__ void __launch_bounds__(1024)
    cell_scan_kernel(const int64_t* __restrict__ n_visible, int n_cells, int32_t* __restrict__ mat,
                     int32_t* __restrict__ cell_total, int32_t* __restrict__ counter, int32_t* __restrict__ cell_offsets,
                     int64_t* __restrict__ n_coarse) {
    __shared__ int s_part[32][33];
    __shared__ int s_warp[32];
    __shared__ bool s_last;
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, row = tid >> 5;
    const int cell = blockIdx.x * 32 + lane;
    const long long n_act = (*n_visible + RK_CHUNK - 1) / RK_CHUNK;
    const long long per = (n_act + 31) / 32;
    const long long k0 = min(row * per, n_act), k1 = min(k0 + per, n_act);
    int sum = 0;
    if (cell < n_cells)
        for (long long k = k0; k < k1; ++k) sum += mat[k * n_cells + cell];
    s_part[row][lane] = sum;
    __syncthreads();
    if (row == 0) {
        int run = 0;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int c = s_part[r][lane];
            s_part[r][lane] = run;
            run += c;
        }
        if (cell < n_cells) cell_total[cell] = run;
    }
    __syncthreads();
    if (cell < n_cells) {
        int run = s_part[row][lane];
        for (long long k = k0; k < k1; ++k) {
            const int c = mat[k * n_cells + cell];
            mat[k * n_cells + cell] = run;
            run += c;
        }
    }
    // the last CTA to get here turns the cell totals into cell offsets
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(counter, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int v = tid < n_cells ? ((const volatile int32_t*)cell_total)[tid] : 0;
    const int mine = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) s_warp[row] = v;
    __syncthreads();
    if (row == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int excl = v - mine + (row ? s_warp[row - 1] : 0);
    if (tid < n_cells) cell_offsets[tid] = excl;
    if (tid == 1023) {
        cell_offsets[n_cells] = excl + mine;
        *n_coarse = excl + mine;
    }
}

__global__ void __launch_bounds__(RK_CHUNK)
    ranked_emit_kernel(int N, long long total, const int32_t* __restrict__ order, const float2* __restrict__ means2d,
                       const int32_t* __restrict__ radii, int tile_size, int tile_w, int tile_h, int cw, int chh, int n_cells,
                       const int32_t* __restrict__ mat, const int32_t* __restrict__ cell_offsets,
                       int32_t* __restrict__ vals) {
    extern __shared__ unsigned s_rk[];
    pdl_wait();
    unsigned* bm = s_rk;                                   // [RK_WARPS][n_cells]: splats (lanes) of warp w touching the cell
    int* pre = (int*)(s_rk + (size_t)RK_WARPS * n_cells);  // [RK_WARPS][n_cells]: position of warp w's first pair in the cell
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long slot0 = (long long)blockIdx.x * RK_CHUNK;
    if (order[slot0] < 0) return;
    for (int i = tid; i < RK_WARPS * n_cells; i += RK_CHUNK) bm[i] = 0;
    __syncthreads();
    const long long slot = slot0 + tid;
    const long long idx = slot < total ? order[slot] : -1;
    int cwid = 0, chgt = 0, n_my = 0, cell0 = 0;  // the splat's cell rectangle: width, height, cells, first cell (+ warp plane)
    if (idx >= 0) {
        const int r = radii[idx];
        if (r > 0) {
            const float2 m = means2d[idx];
            const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h);
            if (t.x1 > t.x0 && t.y1 > t.y0) {
                const int cx0 = t.x0 >> CK_SHIFT, cy0 = t.y0 >> CK_SHIFT;
                cwid = ((t.x1 + CK - 1) >> CK_SHIFT) - cx0;
                chgt = ((t.y1 + CK - 1) >> CK_SHIFT) - cy0;
                n_my = cwid * chgt;
                cell0 = (int)(idx / N) * cw * chh + warp * n_cells + cy0 * cw + cx0;
            }
        }
    }
    const bool big = n_my > RK_BIG;
    const unsigned big_lanes = __ballot_sync(0xffffffffu, big);
    if (!big)
        for (int y = 0, c = cell0; y < chgt; ++y, c += cw)
            for (int x = 0; x < cwid; ++x) atomicOr(bm + c + x, 1u << lane);
    for (unsigned todo = big_lanes; todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const WarpRect R = warp_rect(src, cwid, n_my, cell0);
        for (int i = lane; i < R.n; i += 32) atomicOr(bm + R.cell(i, cw), 1u << src);
    }
    __syncthreads();
    const int32_t* row = mat + (long long)blockIdx.x * n_cells;
    for (int cell = tid; cell < n_cells; cell += RK_CHUNK) {
        int run = cell_offsets[cell] + row[cell];
#pragma unroll
        for (int w = 0; w < RK_WARPS; ++w) {
            pre[w * n_cells + cell] = run;
            run += __popc(bm[w * n_cells + cell]);
        }
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1;
    if (!big)
        for (int y = 0, c = cell0; y < chgt; ++y, c += cw)
            for (int x = 0; x < cwid; ++x) vals[pre[c + x] + __popc(bm[c + x] & lt)] = (int32_t)idx;
    for (unsigned todo = big_lanes; todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const WarpRect R = warp_rect(src, cwid, n_my, cell0);
        const int id = __shfl_sync(0xffffffffu, (int)idx, src);
        const unsigned below = (1u << src) - 1;
        for (int i = lane; i < R.n; i += 32) {
            const int c = R.cell(i, cw);
            vals[pre[c] + __popc(bm[c] & below)] = id;
        }
    }
}

// ---- step 5 -------------------------------------------------------------------------------
// CTA = coarse cell.  Chunks of FB_THREADS list entries (thread = entry); for each of the cell's 16
// tiles the entries that overlap it are ranked with a ballot + per-warp prefix and appended to
// the tile's list.  Running per-tile cursors live in shared memory.  The chunks of a cell are
// sequential (three barriers each), so the CTA is made as wide as possible: with 256 threads the
// kernel time was the longest cell's ~30 chunks (ncu r1_final: 129 us, long-scoreboard bound).
constexpr int FB_THREADS = 1024;
constexpr int FB_WARPS = FB_THREADS / 32;
__global__ void __launch_bounds__(FB_THREADS)
    fine_bin_kernel(int N, const int32_t* __restrict__ coarse_offsets /*[n_cells]*/, long long n_coarse,
                    int n_cells_total, const int32_t* __restrict__ coarse_vals, const float2* __restrict__ means2d,
                    const int32_t* __restrict__ radii, int tile_size, int tile_w, int tile_h, int cw, int chh,
                    const int32_t* __restrict__ isect_offsets, int32_t* __restrict__ flatten_ids) {
    pdl_wait();
    constexpr int NT = CK * CK;
    __shared__ int s_cursor[NT];       // entries already written per tile
    __shared__ int s_warp_cnt[FB_WARPS][NT];  // per chunk: entries per (warp, tile)
    __shared__ int s_tile_off[NT];     // isect_offsets of the cell's tiles (-1 = outside the image)
    const int cell = blockIdx.x;
    const int cam = cell / (cw * chh);
    const int crem = cell - cam * cw * chh;
    const int cy = crem / cw, cx = crem - cy * cw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = coarse_offsets[cell];
    const int end = (cell == n_cells_total - 1) ? (int)n_coarse : coarse_offsets[cell + 1];
    if (end <= start) return;
    if (tid < NT) {
        const int tx = cx * CK + (tid & (CK - 1)), ty = cy * CK + (tid >> CK_SHIFT);
        s_cursor[tid] = 0;
        s_tile_off[tid] = (tx < tile_w && ty < tile_h) ? isect_offsets[(cam * tile_h + ty) * tile_w + tx] : -1;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1;
    // software pipeline: the next chunk's (id, mean, radius) gathers are in flight while this chunk
    // is ranked and written
    int id_n = 0, r_n = 0;
    float2 m_n = make_float2(0.f, 0.f);
    if (start + tid < end) {
        id_n = coarse_vals[start + tid];
        m_n = means2d[id_n];
        r_n = radii[id_n];
    }
    for (int base = start; base < end; base += FB_THREADS) {
        const int e = base + tid;
        const int id = id_n;
        const float2 m = m_n;
        const int r = r_n;
        if (e + FB_THREADS < end) {
            id_n = coarse_vals[e + FB_THREADS];
            m_n = means2d[id_n];
            r_n = radii[id_n];
        }
        unsigned mask = 0;  // bit (iy*CK + ix) set if the splat overlaps tile (cx*CK+ix, cy*CK+iy)
        if (e < end) {
            const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h);
            const int x0 = max(t.x0 - cx * CK, 0), x1 = min(t.x1 - cx * CK, CK);
            const int y0 = max(t.y0 - cy * CK, 0), y1 = min(t.y1 - cy * CK, CK);
            if (x1 > x0 && y1 > y0) {
                const unsigned cols = ((1u << x1) - 1) & ~((1u << x0) - 1);  // CK bits
                const unsigned rows = ((1u << (y1 * CK)) - 1) & ~((1u << (y0 * CK)) - 1);
                mask = (cols * 0x1111u) & rows;  // the column pattern in every row, cut to rows [y0, y1)
            }
        }
        // per-warp counts per tile
        unsigned bal[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            bal[t] = __ballot_sync(0xffffffffu, (mask >> t) & 1u);
            if (lane == 0) s_warp_cnt[warp][t] = __popc(bal[t]);
        }
        __syncthreads();
        // exclusive prefix over warps (threads 0..NT-1), advance the cursors
        if (tid < NT) {
            int run = s_cursor[tid];
#pragma unroll
            for (int w = 0; w < FB_WARPS; ++w) {
                const int c = s_warp_cnt[w][tid];
                s_warp_cnt[w][tid] = run;
                run += c;
            }
            s_cursor[tid] = run;
        }
        __syncthreads();
        const int first = lane < NT ? s_tile_off[lane] + s_warp_cnt[warp][lane] : 0;  // lane t: the warp's first slot in tile t
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int pos = __shfl_sync(0xffffffffu, first, t) + __popc(bal[t] & lt);
            if ((mask >> t) & 1u) flatten_ids[pos] = id;
        }
        __syncthreads();
    }
}


}  // namespace fg

using namespace fg;

extern "C" int fg_bin_coarse_dims(int tile_w, int tile_h, int* cw, int* chh) {
    FG_REQUIRE(cw && chh, "NULL pointer");
    *cw = (tile_w + CK - 1) / CK;
    *chh = (tile_h + CK - 1) / CK;
    return FG_OK;
}

extern "C" int fg_bin_count(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                            int tile_size, int tile_w, int tile_h, int32_t* diff_grid, int32_t* coarse_cnt,
                            void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0 && (long long)C * N < (1ll << 31), "bad C/N");
    FG_REQUIRE(diff_grid, "diff_grid must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    FG_CUDA(cudaMemsetAsync(diff_grid, 0, sizeof(int32_t) * (size_t)C * (tile_h + 1) * (tile_w + 1), st));
    if (N == 0) return FG_OK;
    FG_REQUIRE(order && means2d && radii && coarse_cnt, "NULL pointer");
    const long long total = (long long)C * N;
    FG_LAUNCH(bin_count_kernel, ceil_div(total, 256), 256, 0, st, C, N, total, order, (const float2*)means2d, radii,
              tile_size, tile_w, tile_h, diff_grid, coarse_cnt);
    return FG_OK;
}

extern "C" int fg_exclusive_scan_i32(int64_t n, const int32_t* counts, int32_t* offsets, int64_t* total,
                                     void* workspace, int64_t workspace_bytes, void* stream);
extern "C" int64_t fg_scan_workspace_bytes(int64_t n);

extern "C" int64_t fg_bin_tile_scan_workspace_bytes(int C, int tile_w, int tile_h) {
    const int64_t n = (int64_t)C * tile_w * tile_h;
    return ((n * 4 + 255) & ~(int64_t)255) + fg_scan_workspace_bytes(n);
}

extern "C" int fg_bin_tile_scan(int C, int tile_w, int tile_h, int32_t* diff_grid, int32_t* isect_offsets,
                                int64_t* total, void* workspace, int64_t workspace_bytes, void* stream) {
    FG_REQUIRE(C >= 1 && tile_w > 0 && tile_h > 0 && diff_grid && isect_offsets && total && workspace, "bad arguments");
    FG_REQUIRE(workspace_bytes >= fg_bin_tile_scan_workspace_bytes(C, tile_w, tile_h), "tile-scan workspace too small");
    const int64_t n = (int64_t)C * tile_w * tile_h;
    int32_t* counts = (int32_t*)workspace;
    unsigned char* scan_ws = (unsigned char*)workspace + ((n * 4 + 255) & ~(int64_t)255);
    if (C == 1) {  // one camera = one CTA holds all counts: it scans them too
        FG_LAUNCH(tile_count_kernel, 1, 1024, 0, stream, tile_w, tile_h, diff_grid, counts, isect_offsets, total);
        return FG_OK;
    }
    FG_LAUNCH(tile_count_kernel, C, 1024, 0, stream, tile_w, tile_h, diff_grid, counts, (int32_t*)nullptr, (int64_t*)nullptr);
    return fg_exclusive_scan_i32(n, counts, isect_offsets, total, scan_ws, fg_scan_workspace_bytes(n), stream);
}

extern "C" int fg_bin_coarse_emit(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                                  const int32_t* coarse_off, int tile_size, int tile_w, int tile_h,
                                  uint32_t* coarse_keys, int32_t* coarse_vals, void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0 && (long long)C * N < (1ll << 31), "bad C/N");
    if (N == 0) return FG_OK;
    FG_REQUIRE(order && means2d && radii && coarse_off && coarse_keys && coarse_vals, "NULL pointer");
    const long long total = (long long)C * N;
    const int cw = (tile_w + CK - 1) / CK, chh = (tile_h + CK - 1) / CK;
    FG_LAUNCH(coarse_emit_kernel, ceil_div(total, 256), 256, 0, stream, C, N, total, order, (const float2*)means2d,
              radii, coarse_off, tile_size, tile_w, tile_h, cw, chh, coarse_keys, coarse_vals);
    return FG_OK;
}

extern "C" int64_t fg_bin_ranked_workspace_bytes(int C, int N, int tile_w, int tile_h) {
    if (C < 1 || N < 0 || tile_w < 1 || tile_h < 1) return 0;
    return (int64_t)ranked_layout(C, N, tile_w, tile_h).bytes;
}

extern "C" int fg_bin_count_cells(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                                  int tile_size, int tile_w, int tile_h, int32_t* diff_grid, void* ranked_workspace,
                                  int64_t ranked_workspace_bytes, void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0 && (long long)C * N < (1ll << 31), "bad C/N");
    FG_REQUIRE(diff_grid, "diff_grid must not be NULL");
    const RankedLayout L = ranked_layout(C, N, tile_w, tile_h);
    FG_REQUIRE(L.n_cells > 0, "too many coarse cells for the ranked path (fg_bin_ranked_workspace_bytes == 0)");
    FG_REQUIRE(ranked_workspace && (size_t)ranked_workspace_bytes >= L.bytes, "ranked workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    FG_CUDA(cudaMemsetAsync(diff_grid, 0, sizeof(int32_t) * (size_t)C * (tile_h + 1) * (tile_w + 1), st));
    unsigned char* ws = (unsigned char*)ranked_workspace;
    if (N == 0) {
        FG_CUDA(cudaMemsetAsync(ws + L.counter, 0, 4, st));
        return FG_OK;
    }
    FG_REQUIRE(order && means2d && radii, "NULL pointer");
    const int cw = (tile_w + CK - 1) / CK, chh = (tile_h + CK - 1) / CK;
    FG_LAUNCH(bin_count_cells_kernel, (unsigned)L.n_chunks, RK_CHUNK, (size_t)C * (cw + 1) * (chh + 1) * 4, st, C, N,
              (long long)C * N, order, (const float2*)means2d, radii, tile_size, tile_w, tile_h, cw, chh, L.n_cells, diff_grid,
              (int32_t*)(ws + L.mat), (int32_t*)(ws + L.counter));
    return FG_OK;
}

extern "C" int fg_bin_cell_scan(int C, int N, int tile_w, int tile_h, const int64_t* n_visible, void* ranked_workspace,
                                int64_t ranked_workspace_bytes, int32_t* cell_offsets, int64_t* n_coarse, void* stream) {
    const RankedLayout L = ranked_layout(C, N, tile_w, tile_h);
    FG_REQUIRE(L.n_cells > 0, "too many coarse cells for the ranked path");
    FG_REQUIRE(ranked_workspace && (size_t)ranked_workspace_bytes >= L.bytes, "ranked workspace too small");
    FG_REQUIRE(n_visible && cell_offsets && n_coarse, "NULL pointer");
    unsigned char* ws = (unsigned char*)ranked_workspace;
    FG_LAUNCH(cell_scan_kernel, (L.n_cells + 31) / 32, 1024, 0, stream, n_visible, L.n_cells, (int32_t*)(ws + L.mat),
              (int32_t*)(ws + L.total), (int32_t*)(ws + L.counter), cell_offsets, n_coarse);
    return FG_OK;
}

extern "C" int fg_bin_ranked_emit(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                                  int tile_size, int tile_w, int tile_h, const void* ranked_workspace,
                                  int64_t ranked_workspace_bytes, const int32_t* cell_offsets, int32_t* coarse_vals,
                                  void* stream) {
    const RankedLayout L = ranked_layout(C, N, tile_w, tile_h);
    FG_REQUIRE(L.n_cells > 0, "too many coarse cells for the ranked path");
    FG_REQUIRE(ranked_workspace && (size_t)ranked_workspace_bytes >= L.bytes, "ranked workspace too small");
    if (N == 0) return FG_OK;
    FG_REQUIRE(order && means2d && radii && cell_offsets && coarse_vals, "NULL pointer");
    const int cw = (tile_w + CK - 1) / CK, chh = (tile_h + CK - 1) / CK;
    const size_t smem = (size_t)2 * RK_WARPS * L.n_cells * 4;
    FG_CUDA(cudaFuncSetAttribute(ranked_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned char* ws = (const unsigned char*)ranked_workspace;
    FG_LAUNCH(ranked_emit_kernel, (unsigned)L.n_chunks, RK_CHUNK, smem, stream, N, (long long)C * N, order,
              (const float2*)means2d, radii, tile_size, tile_w, tile_h, cw, chh, L.n_cells, (const int32_t*)(ws + L.mat),
              cell_offsets, coarse_vals);
    return FG_OK;
}

extern "C" int fg_bin_fine(int C, int N, int64_t n_coarse, const int32_t* coarse_offsets,
                           const int32_t* coarse_vals_sorted, const float* means2d, const int32_t* radii,
                           int tile_size, int tile_w, int tile_h, const int32_t* isect_offsets,
                           int32_t* flatten_ids, void* stream) {
    FG_REQUIRE(C >= 1 && n_coarse >= 0 && n_coarse < (1ll << 31), "bad arguments");
    if (n_coarse == 0) return FG_OK;
    FG_REQUIRE(coarse_offsets && coarse_vals_sorted && means2d && radii && isect_offsets && flatten_ids, "NULL pointer");
    const int cw = (tile_w + CK - 1) / CK, chh = (tile_h + CK - 1) / CK;
    const int n_cells = C * cw * chh;
    FG_LAUNCH(fine_bin_kernel, n_cells, FB_THREADS, 0, stream, N, coarse_offsets, (long long)n_coarse, n_cells,
              coarse_vals_sorted, (const float2*)means2d, radii, tile_size, tile_w, tile_h, cw, chh, isect_offsets,
              flatten_ids);
    return FG_OK;
}
