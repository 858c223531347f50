// (2b) Hand-written stable LSD radix sort of (key, uint32 value) pairs -- "onesweep":
// one global histogram pass, then per 8-bit digit ONE kernel that reads every pair once
// and writes it once, chaining the per-digit prefix across tiles with decoupled look-back.
// Replaces cub::DeviceRadixSort::SortPairs inside gsplat's isect_tiles (SURVEY.md 2.2,
// Appendix A.5): a stable sort has exactly one output order, so the result is
// bit-identical to the reference's.
//
// Roofline: HBM.  Per pair: sizeof(key) (histogram read) + P * 2 * (sizeof(key)+4),
// P = ceil(end_bit/8): 152 B at 64-bit keys / 6 passes, 176 B at 7 passes.
//
// Tile = 512 threads x 8 items, warp-striped.  In-tile ranking is the warp
// match-and-count scheme: for each item slot the warp groups equal digits with
// __match_any_sync, the lowest lane of each group bumps that digit's per-warp counter in
// shared memory, and a key's rank is (counter before) + (number of equal-digit lanes below
// it).  Per-warp counters are then prefix-summed across warps and digits, pairs are
// staged in shared memory in tile-sorted order and written out in digit runs.
#include "common.cuh"

namespace fg {

constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 pairs
constexpr int RS_RADIX = 256;

constexpr uint32_t FLAG_LOCAL = 1u << 30;      // tile's own count is published
constexpr uint32_t FLAG_INCLUSIVE = 2u << 30;  // inclusive prefix over tiles 0..t is published
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

// Lanes holding the same 8-bit digit, as a lane mask.  Built from 8 ballots: on sm_100 this is
// several times faster than MATCH.ANY when most lanes hold distinct digits (ncu: MATCH was 38 %
// of all stall samples of the onesweep kernel).  Lanes with valid == false form their own group.
__device__ __forceinline__ uint32_t match_digit(uint32_t d, bool valid, int bits) {
    const uint32_t v = __ballot_sync(0xffffffffu, valid);
    uint32_t peers = valid ? v : ~v;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b >= bits) break;  // warp-uniform: digits of the last pass may be narrower than 8 bits
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// DEPTH = true: the keys are produced here as well -- key = float bits of the depth (positive, so integer order = float
// order), all ones for a splat that touches no tile; vals = the flat index -- and the all-ones keys are left out of the
// histogram (they are dropped by the first sorting pass).
template <typename KeyT, bool DEPTH = false>
__global__ void __launch_bounds__(RS_THREADS)
    rs_histogram_kernel(long long n, const KeyT* __restrict__ keys, int passes, int end_bit,
                        uint32_t* __restrict__ global_hist /*[passes][256]*/, const float* __restrict__ depths = nullptr,
                        const int32_t* __restrict__ tiles_per_gauss = nullptr, KeyT* __restrict__ keys_out = nullptr,
                        uint32_t* __restrict__ vals_out = nullptr) {
    pdl_wait();
    __shared__ uint32_t hist[8 * RS_RADIX];
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS) hist[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * RS_THREADS;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count; lanes past the end vote in a padding group
    for (long long base = (long long)blockIdx.x * RS_THREADS + (threadIdx.x - lane); base < n; base += stride) {
        const long long i = base + lane;
        bool valid = i < n;
        KeyT k = (KeyT)0;
        if (DEPTH) {
            if (valid) {
                k = tiles_per_gauss[i] > 0 ? (KeyT)(uint32_t)__float_as_int(depths[i]) : (KeyT)~(KeyT)0;
                keys_out[i] = k;
                vals_out[i] = (uint32_t)i;
                valid = k != (KeyT)~(KeyT)0;
            }
        } else {
            k = valid ? keys[i] : (KeyT)0;
        }
        // warp-aggregated: lanes with equal digits elect one lane to add their count, so
        // passes whose digit is (nearly) constant do not serialise on one shared-memory word
        for (int p = 0; p < passes; ++p) {
            int shift = 8 * p;
            int bits = min(8, end_bit - shift);
            uint32_t d = (uint32_t)(k >> shift) & ((1u << bits) - 1);
            unsigned peers = match_digit(d, valid, bits);
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[p * RS_RADIX + d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS)
        if (hist[i]) atomicAdd(&global_hist[i], hist[i]);
}

// exclusive scan of each pass's 256 bins, in place; one block of 256 threads per pass
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(uint32_t* __restrict__ global_hist,
                                                                long long* __restrict__ n_sorted = nullptr) {
    __shared__ uint32_t warp_tot[8];
    pdl_wait();
    uint32_t* h = global_hist + blockIdx.x * RS_RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = h[threadIdx.x], incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    h[threadIdx.x] = base + incl - v;
    if (n_sorted && blockIdx.x == 0 && threadIdx.x == RS_RADIX - 1) *n_sorted = base + incl;  // items the passes will move
}

template <typename KeyT>
struct RsSmem {
    KeyT keys[RS_TILE];
    uint32_t vals[RS_TILE];
    uint32_t warp_hist[RS_WARPS * RS_RADIX];
    uint32_t digit_start[RS_RADIX];   // exclusive start of each digit in the tile-sorted order
    int32_t scatter_base[RS_RADIX];   // global position of the digit's first key minus digit_start
    uint32_t warp_tot[RS_RADIX / 32];
    int tile_id;
    int tile_valid;
};

// LB_WIN = status words a look-back round keeps in flight.  Small inputs (every tile resident at
// once, so tile t really has to walk back over t predecessors) use 32; large inputs use 8 to keep
// the register count at 64 (2 CTAs per SM).
// DROP: items whose key is all ones are not sorted at all (the depth sort's first pass drops the culled splats, so the
// remaining passes and everything downstream see the visible ones only).  n_dev != NULL: the item count lives on the device
// (tiles past it exit at once; no tile with work has such a predecessor).
template <typename KeyT, int LB_WIN, bool DROP = false>
__global__ void __launch_bounds__(RS_THREADS)
    rs_onesweep_kernel(long long n, const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                       KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int shift, int bits,
                       const uint32_t* __restrict__ global_offs /*[256] exclusive*/,
                       volatile uint32_t* status /*[tiles][256]*/, int* __restrict__ tile_counter,
                       const long long* __restrict__ n_dev = nullptr) {
    pdl_wait();
    if (n_dev) n = min(n, *n_dev);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem<KeyT>& s = *reinterpret_cast<RsSmem<KeyT>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t digit_mask = (1u << bits) - 1;

    if (tid == 0) s.tile_id = atomicAdd(tile_counter, 1);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) s.warp_hist[i] = 0;
    __syncthreads();
    const int tile = s.tile_id;
    const long long tile_base = (long long)tile * RS_TILE;
    if (tile_base >= n) return;  // (device-side count) block-uniform
    const int tile_n = (int)min((long long)RS_TILE, n - tile_base);

    // ---- load (warp-striped) and rank
    KeyT key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    const int warp_base = warp * (32 * RS_ITEMS);
#pragma unroll
    bool ok[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        int local = warp_base + i * 32 + lane;
        bool valid = local < tile_n;
        key[i] = valid ? keys_in[tile_base + local] : (KeyT)~(KeyT)0;
        val[i] = valid ? vals_in[tile_base + local] : 0u;
        ok[i] = valid && !(DROP && key[i] == (KeyT)~(KeyT)0);
    }
    uint32_t* wh = s.warp_hist + warp * RS_RADIX;
    const uint32_t lt_mask = (1u << lane) - 1;
    // all digit matches first: they are independent, so their latencies overlap
    uint32_t peers[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const bool valid = ok[i];
        const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
        peers[i] = match_digit(d, valid, bits);  // padding lanes: own group
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const bool valid = ok[i];
        const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
        const int leader = __ffs(peers[i]) - 1;
        uint32_t before = 0;
        if (lane == leader && valid) {
            before = wh[d];
            wh[d] = before + __popc(peers[i]);
        }
        before = __shfl_sync(0xffffffffu, before, leader);
        rank[i] = before + __popc(peers[i] & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: exclusive scan over warps, tile total, look-back
    uint32_t tile_count = 0;
    if (tid < RS_RADIX) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = s.warp_hist[w * RS_RADIX + tid];
            s.warp_hist[w * RS_RADIX + tid] = run;
            run += c;
        }
        tile_count = run;
        // publish as early as possible so successors can progress
        volatile uint32_t* my = status + (size_t)tile * RS_RADIX + tid;
        *my = (tile == 0 ? FLAG_INCLUSIVE : FLAG_LOCAL) | tile_count;
        // exclusive scan of tile_count over digits
        uint32_t incl = tile_count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s.warp_tot[warp] = incl;
        s.digit_start[tid] = incl - tile_count;  // completed below with the cross-warp base
    }
    __syncthreads();
    if (tid < RS_RADIX) {
        uint32_t base = 0;
        for (int w = 0; w < warp; ++w) base += s.warp_tot[w];
        uint32_t dstart = s.digit_start[tid] + base;
        // decoupled look-back over predecessor tiles, LB_WIN status words in flight at a time
        uint32_t excl = 0;
        if (tile > 0) {
            int t = tile - 1;
            bool done = false;
            while (!done) {
                uint32_t st[LB_WIN];
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j)
                    st[j] = (t - j >= 0) ? status[(size_t)(t - j) * RS_RADIX + tid] : (2u << 30);  // before tile 0: inclusive, 0
                int used = 0;
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j) {
                    if (done || used != j) continue;  // stop at the first unpublished word; re-poll from there
                    const uint32_t flag = st[j] & FLAG_MASK;
                    if (flag == 0) continue;
                    excl += st[j] & VALUE_MASK;
                    ++used;
                    if (flag == FLAG_INCLUSIVE) done = true;
                }
                t -= used;
            }
            status[(size_t)tile * RS_RADIX + tid] = FLAG_INCLUSIVE | (excl + tile_count);
        }
        s.digit_start[tid] = dstart;
        s.scatter_base[tid] = (int32_t)(global_offs[tid] + excl) - (int32_t)dstart;
        if (tid == RS_RADIX - 1) s.tile_valid = (int)(dstart + tile_count);  // items of this tile that are sorted
    }
    __syncthreads();
    const int tile_valid = DROP ? s.tile_valid : tile_n;

    // ---- stage in tile-sorted order
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        if (ok[i]) {
            uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
            uint32_t pos = s.digit_start[d] + s.warp_hist[warp * RS_RADIX + d] + rank[i];
            s.keys[pos] = key[i];
            s.vals[pos] = val[i];
        }
    }
    __syncthreads();
    // ---- scatter digit runs
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        int pos = i * RS_THREADS + tid;
        if (pos < tile_valid) {
            KeyT k = s.keys[pos];
            uint32_t d = (uint32_t)(k >> shift) & digit_mask;
            long long g = (long long)s.scatter_base[d] + pos;
            keys_out[g] = k;
            vals_out[g] = s.vals[pos];
        }
    }
}

struct RsLayout {
    int passes;
    long long tiles;
    size_t off_hist, off_counters, off_status, total;
};
static RsLayout rs_layout(long long n, int passes) {
    RsLayout L;
    L.passes = passes;
    L.tiles = (n + RS_TILE - 1) / RS_TILE;
    size_t o = 0;
    L.off_hist = o; o += (size_t)8 * RS_RADIX * 4;
    L.off_counters = o; o += 8 * 4; o = (o + 255) & ~(size_t)255;
    L.off_status = o; o += (size_t)passes * L.tiles * RS_RADIX * 4;
    L.total = o;
    return L;
}

template <typename KeyT>
static int radix_sort_pairs(long long n, KeyT* keys_in, uint32_t* vals_in, KeyT* keys_out, uint32_t* vals_out,
                            int end_bit, void* workspace, long long workspace_bytes, int* result_in_out,
                            cudaStream_t st) {
    FG_REQUIRE(n >= 0 && n < (1ll << 30), "n must be in [0, 2^30)");
    FG_REQUIRE(end_bit >= 1 && end_bit <= (int)sizeof(KeyT) * 8, "end_bit out of range");
    FG_REQUIRE(result_in_out, "result_in_out must not be NULL");
    if (n == 0) { *result_in_out = 1; return FG_OK; }
    FG_REQUIRE(keys_in && vals_in && keys_out && vals_out && workspace, "NULL pointer");
    const int passes = (end_bit + 7) / 8;
    RsLayout L = rs_layout(n, passes);
    FG_REQUIRE((size_t)workspace_bytes >= rs_layout(n, 8).total || (size_t)workspace_bytes >= L.total,
               "radix sort workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    uint32_t* hist = (uint32_t*)(ws + L.off_hist);
    int* counters = (int*)(ws + L.off_counters);
    uint32_t* status = (uint32_t*)(ws + L.off_status);
    FG_CUDA(cudaMemsetAsync(ws, 0, L.total, st));
    int hist_blocks = (int)min((long long)num_sms() * 4, (n + RS_THREADS - 1) / RS_THREADS);  // 1 per SM measured slower
    FG_LAUNCH((rs_histogram_kernel<KeyT, false>), hist_blocks, RS_THREADS, 0, st, n, (const KeyT*)keys_in, passes, end_bit, hist,
              (const float*)nullptr, (const int32_t*)nullptr, (KeyT*)nullptr, (uint32_t*)nullptr);
    FG_LAUNCH(rs_scan_hist_kernel, passes, RS_RADIX, 0, st, hist, (long long*)nullptr);
    const size_t smem = sizeof(RsSmem<KeyT>);
    const bool small = false;  // a 32-word window was measured: no gain on 1 M-item sorts (per-tile latency dominates), more registers
    static const cudaError_t attr_once = [] {
        cudaError_t e = cudaFuncSetAttribute(rs_onesweep_kernel<KeyT, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(RsSmem<KeyT>));
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(rs_onesweep_kernel<KeyT, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(RsSmem<KeyT>));
    }();
    FG_CUDA(attr_once);
    KeyT* kin = keys_in; uint32_t* vin = vals_in; KeyT* kout = keys_out; uint32_t* vout = vals_out;
    for (int p = 0; p < passes; ++p) {
        int shift = 8 * p;
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        if (small) {
            FG_LAUNCH((rs_onesweep_kernel<KeyT, 32>), (int)L.tiles, RS_THREADS, smem, st, n, (const KeyT*)kin, (const uint32_t*)vin, kout, vout, shift,
                      bits, (const uint32_t*)(hist + p * RS_RADIX), status + (size_t)p * L.tiles * RS_RADIX, counters + p,
                      (const long long*)nullptr);
        } else {
            FG_LAUNCH((rs_onesweep_kernel<KeyT, 8>), (int)L.tiles, RS_THREADS, smem, st, n, (const KeyT*)kin, (const uint32_t*)vin, kout, vout, shift,
                          bits, (const uint32_t*)(hist + p * RS_RADIX), status + (size_t)p * L.tiles * RS_RADIX, counters + p,
                          (const long long*)nullptr);
        }
        KeyT* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *result_in_out = (kin == keys_out) ? 1 : 0;  // after the last swap `kin` holds the result
    return FG_OK;
}

}  // namespace fg

using namespace fg;

static inline size_t rs_al(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int64_t fg_depth_sort_workspace_bytes(int64_t total) {
    const long long n = total < 1 ? 1 : total;
    return (int64_t)(4 * rs_al((size_t)n * 4) + rs_layout(n, 4).total);
}

// Depth order of the VISIBLE splats in one call: keys (+ the four digit histograms) from depths / tiles_per_gauss, a first
// pass that drops the culled splats while it sorts, three passes over the visible ones only.
extern "C" int fg_depth_sort_visible(int64_t total, const float* depths, const int32_t* tiles_per_gauss, int32_t* order,
                                     int64_t* n_visible_dev, void* workspace, int64_t workspace_bytes, void* stream) {
    FG_REQUIRE(total >= 0 && total < (1ll << 30), "total must be in [0, 2^30)");
    FG_REQUIRE(n_visible_dev != nullptr, "n_visible_dev must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (total == 0) {
        FG_CUDA(cudaMemsetAsync(n_visible_dev, 0, 8, st));
        return FG_OK;
    }
    FG_REQUIRE(depths && tiles_per_gauss && order && workspace, "NULL pointer");
    FG_REQUIRE(workspace_bytes >= fg_depth_sort_workspace_bytes(total), "depth-sort workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    const size_t arr = rs_al((size_t)total * 4);
    uint32_t* kA = (uint32_t*)ws; uint32_t* vA = (uint32_t*)(ws + arr);
    uint32_t* kB = (uint32_t*)(ws + 2 * arr); uint32_t* vB = (uint32_t*)(ws + 3 * arr);
    unsigned char* rs = ws + 4 * arr;
    const RsLayout L = rs_layout(total, 4);
    uint32_t* hist = (uint32_t*)(rs + L.off_hist);
    int* counters = (int*)(rs + L.off_counters);
    uint32_t* status = (uint32_t*)(rs + L.off_status);
    FG_CUDA(cudaMemsetAsync(rs, 0, L.total, st));
    FG_CUDA(cudaMemsetAsync(order, 0xff, (size_t)total * 4, st));  // -1 past the visible splats: the consumers stop there
    const int hist_blocks = (int)min((long long)num_sms() * 4, ((long long)total + RS_THREADS - 1) / RS_THREADS);
    FG_LAUNCH((rs_histogram_kernel<uint32_t, true>), hist_blocks, RS_THREADS, 0, st, (long long)total, (const uint32_t*)nullptr, 4, 32,
              hist, depths, tiles_per_gauss, kA, vA);
    FG_LAUNCH(rs_scan_hist_kernel, 4, RS_RADIX, 0, st, hist, (long long*)n_visible_dev);
    const size_t smem = sizeof(RsSmem<uint32_t>);
    static const cudaError_t attr_once = [] {
        cudaError_t e = cudaFuncSetAttribute(rs_onesweep_kernel<uint32_t, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(RsSmem<uint32_t>));
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(rs_onesweep_kernel<uint32_t, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(RsSmem<uint32_t>));
    }();
    FG_CUDA(attr_once);
    const long long* nv = (const long long*)n_visible_dev;
    // pass 0 reads every splat and keeps the visible ones; passes 1-3 are launched for the worst case and read the count on
    // the device (tiles past it exit at once); the last pass writes its values straight into `order`
    FG_LAUNCH((rs_onesweep_kernel<uint32_t, 8, true>), (int)L.tiles, RS_THREADS, smem, st, (long long)total, (const uint32_t*)kA,
              (const uint32_t*)vA, kB, vB, 0, 8, (const uint32_t*)hist, status, counters, (const long long*)nullptr);
    FG_LAUNCH((rs_onesweep_kernel<uint32_t, 8, false>), (int)L.tiles, RS_THREADS, smem, st, (long long)total, (const uint32_t*)kB,
              (const uint32_t*)vB, kA, vA, 8, 8, (const uint32_t*)(hist + RS_RADIX), status + (size_t)L.tiles * RS_RADIX, counters + 1, nv);
    FG_LAUNCH((rs_onesweep_kernel<uint32_t, 8, false>), (int)L.tiles, RS_THREADS, smem, st, (long long)total, (const uint32_t*)kA,
              (const uint32_t*)vA, kB, vB, 16, 8, (const uint32_t*)(hist + 2 * RS_RADIX), status + (size_t)2 * L.tiles * RS_RADIX, counters + 2, nv);
    FG_LAUNCH((rs_onesweep_kernel<uint32_t, 8, false>), (int)L.tiles, RS_THREADS, smem, st, (long long)total, (const uint32_t*)kB,
              (const uint32_t*)vB, kA, (uint32_t*)order, 24, 8, (const uint32_t*)(hist + 3 * RS_RADIX), status + (size_t)3 * L.tiles * RS_RADIX, counters + 3, nv);
    return FG_OK;
}

extern "C" int64_t fg_radix_sort_workspace_bytes(int64_t n) { return (int64_t)rs_layout(n < 1 ? 1 : n, 8).total; }

extern "C" int fg_radix_sort_pairs_u64_u32(int64_t n, uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out,
                                           uint32_t* vals_out, int end_bit, void* workspace,
                                           int64_t workspace_bytes, int* result_in_out, void* stream) {
    return radix_sort_pairs<unsigned long long>(n, (unsigned long long*)keys_in, vals_in,
                                                (unsigned long long*)keys_out, vals_out, end_bit, workspace,
                                                workspace_bytes, result_in_out, (cudaStream_t)stream);
}

extern "C" int fg_radix_sort_pairs_u32_u32(int64_t n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                                           uint32_t* vals_out, int end_bit, void* workspace,
                                           int64_t workspace_bytes, int* result_in_out, void* stream) {
    return radix_sort_pairs<uint32_t>(n, keys_in, vals_in, keys_out, vals_out, end_bit, workspace, workspace_bytes,
                                      result_in_out, (cudaStream_t)stream);
}
