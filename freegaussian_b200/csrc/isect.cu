// (2a) Tile-intersection bookkeeping: exclusive scan of per-splat tile counts, emission
// of (camera|tile|depth) keys, and per-tile offsets into the sorted list.
// Replaces gsplat isect_tiles (+torch.cumsum) and isect_offset_encode
// (tile_size=16 at freegaussian_model.py:806).
//
// Roofline: HBM.  Scan: 8 B read + 4 B written per (c,n).  Emission: 20 B read per (c,n),
// 12 B written per intersection.  Offsets: 8 B read per intersection + 4 B per tile.
#include "common.cuh"
#include "splat_math.h"

namespace fg {

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 counts per block

__device__ __forceinline__ long long warp_incl_scan(long long v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// block-wide exclusive scan of one int64 per thread; returns exclusive prefix, total via *total
__device__ __forceinline__ long long block_excl_scan(long long v, long long* total, long long* warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < (SCAN_THREADS / 32) ? warp_sums[lane] : 0;
        long long wi = warp_incl_scan(w);
        if (lane < (SCAN_THREADS / 32)) warp_sums[lane] = wi - w;
        if (lane == 31) warp_sums[32] = wi;
    }
    __syncthreads();
    long long excl = incl - v + warp_sums[warp];
    *total = warp_sums[32];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(long long n, const int32_t* __restrict__ counts,
                                                                   long long* __restrict__ block_sums) {
    pdl_wait();
    __shared__ long long warp_sums[33];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        long long idx = base + i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += counts[idx];
    }
    long long total;
    block_excl_scan(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place, grand total -> *total
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(int nblocks, long long* __restrict__ block_sums,
                                                                  long long* __restrict__ total_out) {
    pdl_wait();
    __shared__ long long warp_sums[33];
    long long carry = 0;
    for (int base = 0; base < nblocks; base += SCAN_THREADS) {
        int i = base + threadIdx.x;
        long long v = i < nblocks ? block_sums[i] : 0;
        long long total;
        long long excl = block_excl_scan(v, &total, warp_sums);
        if (i < nblocks) block_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(long long n, const int32_t* __restrict__ counts,
                                                                  const long long* __restrict__ block_sums,
                                                                  int32_t* __restrict__ offsets) {
    pdl_wait();
    __shared__ long long warp_sums[33];
    // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? counts[base + i] : 0;
        s += v[i];
    }
    long long total;
    long long run = block_excl_scan(s, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) offsets[base + i] = (int32_t)run;
        run += v[i];
    }
}

// ------------------------------------------------------------------ emission
// One thread per (c,n).  Splats touching many tiles are emitted cooperatively by their
// whole warp so a single huge splat does not serialise one lane.
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_COOP_MIN = 32;  // tiles; at or above this the warp shares the work

// KEY64: key = cam << (32+tile_bits) | tile << 32 | depth bits (the reference layout), splats
// visited in flattened (c*N+n) order.  !KEY64 (two-level path): splats visited in the order
// given by `order` (depth-sorted flat ids) and key = cam*tiles_per_cam + tile (uint32).
template <bool KEY64>
__global__ void __launch_bounds__(EMIT_THREADS)
    isect_emit_kernel(int C, int N, long long total, const int32_t* __restrict__ order,
                      const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                      const float* __restrict__ depths, const int32_t* __restrict__ offsets, int tile_size,
                      int tile_w, int tile_h, int tile_bits, int64_t* __restrict__ isect_ids,
                      uint32_t* __restrict__ tile_keys, int32_t* __restrict__ flatten_ids) {
    pdl_wait();
    const long long slot = (long long)blockIdx.x * EMIT_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, cnt = 0, off = 0;
    int64_t key_base = 0;
    long long idx = -1;
    if (slot < total) {
        idx = KEY64 ? slot : (long long)order[slot];
        int r = radii[idx];
        if (r > 0) {
            float2 m = means2d[idx];
            TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h);
            x0 = t.x0; x1 = t.x1; y0 = t.y0; y1 = t.y1;
            cnt = (x1 - x0) * (y1 - y0);
            off = offsets[slot];
            int64_t cam = idx / N;
            if (KEY64) key_base = (cam << (32 + tile_bits)) | (int64_t)(uint32_t)__float_as_int(depths[idx]);
            else key_base = cam * (int64_t)(tile_w * tile_h);
        }
    }
    // small splats: each lane writes its own
    if (cnt > 0 && cnt < EMIT_COOP_MIN) {
        int k = off;
        for (int i = y0; i < y1; ++i)
            for (int j = x0; j < x1; ++j) {
                if (KEY64) isect_ids[k] = key_base | ((int64_t)(i * tile_w + j) << 32);
                else tile_keys[k] = (uint32_t)(key_base + i * tile_w + j);
                flatten_ids[k] = (int32_t)idx;
                ++k;
            }
    }
    // large splats: whole warp cooperates, one splat at a time
    unsigned big = __ballot_sync(0xffffffffu, cnt >= EMIT_COOP_MIN);
    while (big) {
        int src = __ffs(big) - 1;
        big &= big - 1;
        int bx0 = __shfl_sync(0xffffffffu, x0, src), bx1 = __shfl_sync(0xffffffffu, x1, src);
        int by0 = __shfl_sync(0xffffffffu, y0, src);
        int bcnt = __shfl_sync(0xffffffffu, cnt, src), boff = __shfl_sync(0xffffffffu, off, src);
        long long bkey = __shfl_sync(0xffffffffu, (long long)key_base, src);
        long long bidx = __shfl_sync(0xffffffffu, idx, src);
        int w = bx1 - bx0;
        for (int t = lane; t < bcnt; t += 32) {
            int i = by0 + t / w, j = bx0 + t % w;
            if (KEY64) isect_ids[boff + t] = (int64_t)bkey | ((int64_t)(i * tile_w + j) << 32);
            else tile_keys[boff + t] = (uint32_t)(bkey + i * tile_w + j);
            flatten_ids[boff + t] = (int32_t)bidx;
        }
    }
}

// ------------------------------------------------------------------ offsets
// offsets[t] = first index whose (cam,tile) id is >= t.  Thread i compares id(i-1), id(i)
// and fills every tile id in (id(i-1), id(i)]; the last thread also fills the tail.
template <bool KEY64>
__global__ void __launch_bounds__(256)
    isect_offsets_kernel(long long n_isects, const int64_t* __restrict__ sorted_ids,
                         const uint32_t* __restrict__ sorted_tile_keys, int n_tiles_per_cam, int tile_bits,
                         long long n_tiles_total, int32_t* __restrict__ offsets) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_isects) return;
    auto lin = [&](long long j) -> long long {
        if (!KEY64) return (long long)sorted_tile_keys[j];
        long long id = sorted_ids[j] >> 32;
        long long cam = id >> tile_bits;
        long long tile = id & ((1ll << tile_bits) - 1);
        return cam * n_tiles_per_cam + tile;
    };
    const long long cur = lin(i);
    const long long prev = (i == 0) ? -1 : lin(i - 1);
    for (long long t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
    if (i == n_isects - 1)
        for (long long t = cur + 1; t < n_tiles_total; ++t) offsets[t] = (int32_t)n_isects;
}

// two-level path helpers ------------------------------------------------------------------
// depth key of every (c,n): float bits of the depth (positive, so integer order = float order);
// splats that touch no tile sort to the end.
__global__ void depth_keys_kernel(long long total, const float* __restrict__ depths,
                                  const int32_t* __restrict__ tiles_per_gauss, uint32_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    keys[i] = tiles_per_gauss[i] > 0 ? (uint32_t)__float_as_int(depths[i]) : 0xffffffffu;
    vals[i] = (uint32_t)i;
}
__global__ void gather_i32_kernel(long long n, const int32_t* __restrict__ src, const int32_t* __restrict__ idx,
                                  int32_t* __restrict__ dst) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
// rebuild the reference's 64-bit keys from the two-level result (meta["isect_ids"], on demand)
__global__ void isect_ids_kernel(long long n, const uint32_t* __restrict__ tile_keys,
                                 const int32_t* __restrict__ flatten_ids, const float* __restrict__ depths,
                                 int n_tiles_per_cam, int tile_bits, int64_t* __restrict__ out) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long lin = tile_keys[i];
    long long cam = lin / n_tiles_per_cam, tile = lin - cam * n_tiles_per_cam;
    out[i] = (cam << (32 + tile_bits)) | (tile << 32) | (long long)(uint32_t)__float_as_int(depths[flatten_ids[i]]);
}

// densification statistics (freegaussian_model.py:369-392), all of this rank's views in one pass
__global__ void densify_stats_kernel(int C, int N, const int32_t* __restrict__ radii, const float2* __restrict__ absgrad,
                                     float inv_max_hw, float* __restrict__ grad_norm, float* __restrict__ vis_count,
                                     float* __restrict__ max_size) {
    pdl_wait();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float g = 0.f, cnt = 0.f;
    int rmax = 0;
    for (int c = 0; c < C; ++c) {
        const size_t i = (size_t)c * N + n;
        const int r = radii[i];
        if (r > 0) {
            const float2 a = absgrad[i];
            g += sqrtf(a.x * a.x + a.y * a.y);
            cnt += 1.f;
            rmax = max(rmax, r);
        }
    }
    if (cnt > 0.f) {
        grad_norm[n] += g;
        vis_count[n] += cnt;
        max_size[n] = fmaxf(max_size[n], (float)rmax * inv_max_hw);
    }
}

__global__ void fill_i32_kernel(long long n, int32_t v, int32_t* out) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_scan_workspace_bytes(int64_t n) {
    int64_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (nblocks + 1) * (int64_t)sizeof(long long);
}

extern "C" int fg_exclusive_scan_i32(int64_t n, const int32_t* counts, int32_t* offsets, int64_t* total,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
    FG_REQUIRE(n >= 0 && total, "n must be >= 0 and total must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        FG_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return FG_OK;
    }
    FG_REQUIRE(counts && offsets && workspace, "counts/offsets/workspace must not be NULL");
    FG_REQUIRE(workspace_bytes >= fg_scan_workspace_bytes(n), "scan workspace too small");
    int nblocks = ceil_div(n, SCAN_TILE);
    long long* block_sums = (long long*)workspace;
    FG_LAUNCH(scan_reduce_kernel, nblocks, SCAN_THREADS, 0, st, (long long)n, counts, block_sums);
    FG_LAUNCH(scan_spine_kernel, 1, SCAN_THREADS, 0, st, nblocks, block_sums, (long long*)total);
    FG_LAUNCH(scan_apply_kernel, nblocks, SCAN_THREADS, 0, st, (long long)n, counts, block_sums, offsets);
    return FG_OK;
}

extern "C" int fg_isect_emit(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                             const int32_t* offsets, int tile_size, int tile_w, int tile_h, int64_t* isect_ids,
                             int32_t* flatten_ids, void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0 && (long long)C * N < (1ll << 31), "bad C/N");
    FG_REQUIRE(tile_size > 0 && tile_w > 0 && tile_h > 0, "bad tile geometry");
    if (N == 0) return FG_OK;
    FG_REQUIRE(means2d && radii && depths && offsets && isect_ids && flatten_ids, "NULL pointer");
    int tile_bits = tile_bits_of(tile_w * tile_h);
    long long total = (long long)C * N;
    FG_LAUNCH((isect_emit_kernel<true>), ceil_div(total, EMIT_THREADS), EMIT_THREADS, 0, stream, C, N, total,
              (const int32_t*)nullptr, (const float2*)means2d, radii, depths, offsets, tile_size, tile_w, tile_h,
              tile_bits, isect_ids, (uint32_t*)nullptr, flatten_ids);
    return FG_OK;
}

extern "C" int fg_isect_depth_keys(int64_t total, const float* depths, const int32_t* tiles_per_gauss,
                                   uint32_t* keys, uint32_t* vals, void* stream) {
    FG_REQUIRE(total >= 0 && total < (1ll << 31), "total out of range");
    if (total == 0) return FG_OK;
    FG_REQUIRE(depths && tiles_per_gauss && keys && vals, "NULL pointer");
    FG_LAUNCH(depth_keys_kernel, ceil_div(total, 256), 256, 0, stream, (long long)total, depths, tiles_per_gauss,
              keys, vals);
    return FG_OK;
}

extern "C" int fg_gather_i32(int64_t n, const int32_t* src, const int32_t* idx, int32_t* dst, void* stream) {
    FG_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return FG_OK;
    FG_REQUIRE(src && idx && dst, "NULL pointer");
    FG_LAUNCH(gather_i32_kernel, ceil_div(n, 256), 256, 0, stream, (long long)n, src, idx, dst);
    return FG_OK;
}

extern "C" int fg_isect_emit_tiles(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                                   const int32_t* offsets, int tile_size, int tile_w, int tile_h,
                                   uint32_t* tile_keys, int32_t* flatten_ids, void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0 && (long long)C * N < (1ll << 31), "bad C/N");
    FG_REQUIRE(tile_size > 0 && tile_w > 0 && tile_h > 0, "bad tile geometry");
    FG_REQUIRE((long long)C * tile_w * tile_h < (1ll << 32), "too many tiles for 32-bit tile keys");
    if (N == 0) return FG_OK;
    FG_REQUIRE(order && means2d && radii && offsets && tile_keys && flatten_ids, "NULL pointer");
    long long total = (long long)C * N;
    FG_LAUNCH((isect_emit_kernel<false>), ceil_div(total, EMIT_THREADS), EMIT_THREADS, 0, stream, C, N, total, order,
              (const float2*)means2d, radii, (const float*)nullptr, offsets, tile_size, tile_w, tile_h, 0,
              (int64_t*)nullptr, tile_keys, flatten_ids);
    return FG_OK;
}

extern "C" int fg_isect_offsets_tiles(int64_t n_isects, const uint32_t* sorted_tile_keys, int C, int tile_w,
                                      int tile_h, int32_t* offsets, void* stream) {
    FG_REQUIRE(n_isects >= 0 && n_isects < (1ll << 31), "n_isects must be in [0, 2^31)");
    FG_REQUIRE(C >= 1 && tile_w > 0 && tile_h > 0 && offsets, "bad arguments");
    long long n_tiles = (long long)C * tile_w * tile_h;
    if (n_isects == 0) {
        FG_LAUNCH(fill_i32_kernel, ceil_div(n_tiles, 256), 256, 0, stream, n_tiles, 0, offsets);
        return FG_OK;
    }
    FG_REQUIRE(sorted_tile_keys, "sorted_tile_keys must not be NULL");
    FG_LAUNCH((isect_offsets_kernel<false>), ceil_div(n_isects, 256), 256, 0, stream, (long long)n_isects,
              (const int64_t*)nullptr, sorted_tile_keys, tile_w * tile_h, 0, n_tiles, offsets);
    return FG_OK;
}

extern "C" int fg_isect_ids_from_tiles(int64_t n_isects, const uint32_t* sorted_tile_keys,
                                       const int32_t* flatten_ids, const float* depths, int tile_w, int tile_h,
                                       int64_t* isect_ids, void* stream) {
    FG_REQUIRE(n_isects >= 0, "n_isects must be >= 0");
    if (n_isects == 0) return FG_OK;
    FG_REQUIRE(sorted_tile_keys && flatten_ids && depths && isect_ids, "NULL pointer");
    FG_LAUNCH(isect_ids_kernel, ceil_div(n_isects, 256), 256, 0, stream, (long long)n_isects, sorted_tile_keys,
              flatten_ids, depths, tile_w * tile_h, tile_bits_of(tile_w * tile_h), isect_ids);
    return FG_OK;
}

extern "C" int fg_isect_offsets(int64_t n_isects, const int64_t* sorted_isect_ids, int C, int tile_w, int tile_h,
                                int32_t* offsets, void* stream) {
    FG_REQUIRE(n_isects >= 0 && n_isects < (1ll << 31), "n_isects must be in [0, 2^31)");
    FG_REQUIRE(C >= 1 && tile_w > 0 && tile_h > 0 && offsets, "bad arguments");
    long long n_tiles = (long long)C * tile_w * tile_h;
    if (n_isects == 0) {
        FG_LAUNCH(fill_i32_kernel, ceil_div(n_tiles, 256), 256, 0, stream, n_tiles, 0, offsets);
        return FG_OK;
    }
    FG_REQUIRE(sorted_isect_ids, "sorted_isect_ids must not be NULL");
    int tile_bits = tile_bits_of(tile_w * tile_h);
    FG_LAUNCH((isect_offsets_kernel<true>), ceil_div(n_isects, 256), 256, 0, stream, (long long)n_isects,
              sorted_isect_ids, (const uint32_t*)nullptr, tile_w * tile_h, tile_bits, n_tiles, offsets);
    return FG_OK;
}

extern "C" int fg_densify_stats(int C, int N, const int32_t* radii, const float* absgrad, float inv_max_hw,
                                float* grad_norm, float* vis_count, float* max_size, void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0, "bad C/N");
    if (N == 0) return FG_OK;
    FG_REQUIRE(radii && absgrad && grad_norm && vis_count && max_size, "NULL pointer");
    FG_LAUNCH(densify_stats_kernel, ceil_div(N, 256), 256, 0, stream, C, N, radii, (const float2*)absgrad, inv_max_hw,
              grad_norm, vis_count, max_size);
    return FG_OK;
}
