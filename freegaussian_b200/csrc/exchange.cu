// (5) View-sharded multi-GPU gradient exchange over NVLink peer memory / NVSwitch multicast
// (BASELINE.json north_star item 5).  The reference has no multi-GPU code; what is
// exchanged is defined by the single-process semantics: every rank must end the backward pass holding
// the gradient a single process rendering ALL views would have computed.
//
// Design.  The dense parameter gradient is 236 B per Gaussian, 81 % of it the
// spherical-harmonics rows.  An SH row's gradient is rank one per (view, Gaussian):
//     v_sh[n, k, :] = sum over views c of  basis_k(dir(n, c)) * v_rgb[c, n, :]
// so instead of all-reducing 192 B per Gaussian, every rank PUBLISHES the 12-byte colour gradient of its
// own views (clamp mask applied, plus a 1-bit-per-Gaussian visibility mask) in symmetric memory, and
// `sh_bwd_views_kernel` on every rank reads all ranks' published rows straight over NVLink (peer loads,
// rows of invisible splats are never fetched), evaluates the basis for every view and writes the summed
// rows once: the transfer overlaps the math row by row, the sum order is fixed (rank, view), hence the
// result is bit-identical on every rank.  Only the geometry part (means, quats, scales, opacity, frame
// t+1 means: 56 B per Gaussian) needs a reduction: `allreduce_kernel` is a two-shot all-reduce that
// runs in the switch -- each rank `multimem.ld_reduce`s its 1/G slice (NVSwitch sums the G copies in
// flight) and `multimem.st`s the result to all ranks -- with the cross-rank barriers inside the kernel
// (release/acquire flags in symmetric memory).  Without multicast the same kernel falls back to peer
// loads + peer stores.
#include <stdio.h>

#include <algorithm>

#include "common.cuh"
#include "splat_math.h"

namespace fg {

constexpr int XB = 1024;             // most threads per block of the all-reduce kernel (one such CTA fills an SM)
constexpr int XCHG_SLOT_WORDS = 16;  // one flag word per source rank (FG_XCHG_MAX_RANKS)

struct Peers {
    int world, rank;
    char* buf[FG_XCHG_MAX_RANKS];
    char* mc;
    uint32_t* flags[FG_XCHG_MAX_RANKS];
};

__device__ __forceinline__ void flag_store_release(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t flag_load_acquire(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Barrier among the blocks with the same index on every rank.  Thread r (< world) tells rank r "block
// `slot` of rank `me` reached epoch e" and waits for the same word from rank r.  Epochs only grow, so
// the flags are never reset.  Bounded: a peer that never arrives traps after ~20 s instead of hanging.
__device__ __forceinline__ void block_barrier(const Peers& p, int slot, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x < p.world) {
        __threadfence_system();
        const int r = threadIdx.x;
        flag_store_release(p.flags[r] + slot * XCHG_SLOT_WORDS + p.rank, epoch);
        const uint32_t* mine = p.flags[p.rank] + slot * XCHG_SLOT_WORDS + r;
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(flag_load_acquire(mine) - epoch) < 0) {
            if (globaltimer_ns() - t0 > 20000000000ull) {
                printf("fg exchange: rank %d timed out waiting for rank %d (slot %d, epoch %u)\n", p.rank, r, slot, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(32) xchg_barrier_kernel(Peers p, uint32_t epoch) {
    pdl_wait();
    block_barrier(p, 0, epoch);
}

__device__ __forceinline__ float4 mc_ld_reduce(const float4* a) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(a)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* a, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// In-place SUM over ranks of n4 float4 at byte offset `off` of the symmetric buffer.  Two-shot: rank r
// owns the r-th slice.  Slots 1..grid: start barrier (optional), grid+1..2*grid: end barrier.
// Few, fat CTAs (default 64 x 1024 threads, eight 16-byte requests per thread = 8 MB in flight: enough for the switch round
// trip at link rate) so that the rest of the GPU stays free for the kernels that run beside it -- a first version
// with two CTAs on every SM measured 0.17 ms alone but starved the SH-row kernel of registers: no overlap at all.
template <bool MC>
__global__ void __launch_bounds__(XB, 1) allreduce_kernel(Peers p, long long off, long long n4, uint32_t epoch,
                                                       int start_barrier) {
    pdl_wait();
    if (start_barrier) block_barrier(p, 1 + blockIdx.x, epoch);
    const long long per = (n4 + p.world - 1) / p.world;
    const long long lo = per * p.rank, hi = min(n4, lo + per);
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = MC ? 8 : 4;
    if (MC) {
        float4* mc = reinterpret_cast<float4*>(p.mc + off);
        long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + (U - 1) * stride < hi; i += U * stride) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = mc_ld_reduce(mc + i + u * stride);
#pragma unroll
            for (int u = 0; u < U; ++u) mc_st(mc + i + u * stride, v[u]);
        }
        for (; i < hi; i += stride) mc_st(mc + i, mc_ld_reduce(mc + i));
    } else {
        long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + (U - 1) * stride < hi; i += U * stride) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < p.world; ++r) {  // fixed order: the sum is the same on every rank
                const float4* src = reinterpret_cast<const float4*>(p.buf[r] + off);
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = src[i + u * stride];
#pragma unroll
                for (int u = 0; u < U; ++u) { s[u].x += v[u].x; s[u].y += v[u].y; s[u].z += v[u].z; s[u].w += v[u].w; }
            }
            for (int r = 0; r < p.world; ++r) {
                float4* dst = reinterpret_cast<float4*>(p.buf[r] + off);
#pragma unroll
                for (int u = 0; u < U; ++u) dst[i + u * stride] = s[u];
            }
        }
        for (; i < hi; i += stride) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < p.world; ++r) {
                const float4 v = *(reinterpret_cast<const float4*>(p.buf[r] + off) + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            for (int r = 0; r < p.world; ++r) *(reinterpret_cast<float4*>(p.buf[r] + off) + i) = s;
        }
    }
    block_barrier(p, 1 + gridDim.x + blockIdx.x, epoch);
}

// ------------------------------------------------------------------------------ SH rows from all views
// Published block of one rank (byte offset `pub_off` of its symmetric buffer), V views per rank, `words` = ceil(N/32)
// rounded up to 4:
//   float    campos[V][4]            camera centres (world); then one uint32: nnz = visible (view, Gaussian) pairs
//   uint32_t mask[V][words]          bit n of view v = splat n is visible in the view
//   uint32_t prefix[V][words]        compact row of the first visible splat of that word (exclusive scan over v*N+n)
//   float    rgb[nnz][3]             d loss / d rgb of the visible pairs (clamp mask applied), in ascending v*N+n order
// Everything a peer needs is one contiguous range of fixed_bytes + 12 nnz bytes: fine-grained remote loads are latency
// bound (8 views x 2 round trips per Gaussian cost 0.23 ms at 8 ranks), so peers PULL that range with coalesced
// 16-byte loads (xchg_pull_kernel) and the SH rows are then rebuilt from local memory.
struct PubLayout {
    long long nnz_off, mask_off, prefix_off, rgb_off;  // byte offsets inside the block
    int words;
};
__host__ __device__ inline PubLayout pub_layout(int V, int N) {
    PubLayout L;
    L.words = ((N + 31) / 32 + 3) & ~3;
    L.nnz_off = (long long)V * 16;
    L.mask_off = (L.nnz_off + 16 + 255) / 256 * 256;
    L.prefix_off = L.mask_off + (long long)V * L.words * 4;
    L.rgb_off = (L.prefix_off + (long long)V * L.words * 4 + 255) / 256 * 256;
    return L;
}

struct ViewSrc {
    const char* base[FG_XCHG_MAX_RANKS];  // published block of every rank as THIS rank reads it (own block / pulled copy)
};

// Copy every peer's published range into local staging memory.  grid = (blocks per peer, world).
__global__ void __launch_bounds__(512) xchg_pull_kernel(Peers p, long long pub_off, long long nnz_off, long long fixed_bytes,
                                                         char* staging, long long staging_stride) {
    pdl_wait();
    const int r = blockIdx.y;
    if (r == p.rank) return;
    const char* src = p.buf[r] + pub_off;
    char* dst = staging + (long long)r * staging_stride;
    const unsigned nnz = *reinterpret_cast<const unsigned*>(src + nnz_off);
    const long long n16 = (fixed_bytes + (long long)nnz * 12 + 15) / 16;
    const uint4* s16 = reinterpret_cast<const uint4*>(src);
    uint4* d16 = reinterpret_cast<uint4*>(dst);
    const long long stride = (long long)gridDim.x * 512;
    long long i = (long long)blockIdx.x * 512 + threadIdx.x;
    constexpr int U = 8;
    for (; i + (U - 1) * stride < n16; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = s16[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) d16[i + u * stride] = v[u];
    }
    for (; i < n16; i += stride) d16[i] = s16[i];
}

constexpr int VB = 128;  // Gaussians per block

template <int DEG>
__global__ void __launch_bounds__(VB) sh_bwd_views_kernel(ViewSrc vs, int world, int V, int N,
                                                          const float* __restrict__ means, float* __restrict__ v_sh,
                                                          int sh_row_floats) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int NG = (K + 3) / 4;
    constexpr int NV3 = NG * 3;
    constexpr int ROW = 13;  // odd float4 stride: conflict-free own-row access
    float4* acc = reinterpret_cast<float4*>(smem) + threadIdx.x * ROW;
    float* campos = smem + VB * ROW * 4;  // [world*V][4]
    const PubLayout L = pub_layout(V, N);
    const int n0 = blockIdx.x * VB, n = n0 + threadIdx.x;
    const bool in_range = n < N;
    for (int i = threadIdx.x; i < world * V * 4; i += VB) {
        const int r = i / (V * 4);
        campos[i] = reinterpret_cast<const float*>(vs.base[r])[i - r * V * 4];
    }
#pragma unroll
    for (int j = 0; j < NV3; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (in_range) {
        float m[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i] = __ldg(means + 3 * (size_t)n + i);
        // Views are taken eight at a time: first the eight visibility words and prefixes, then the colour gradients of
        // the views that see this splat -- all loads of a stage are in flight together.  The accumulation order stays
        // (rank, view).
        const int T = world * V;
        constexpr int VC = 8;
        for (int base = 0; base < T; base += VC) {
            uint32_t bits = 0;
            const float* src[VC];
#pragma unroll
            for (int j = 0; j < VC; ++j) {
                const int view = base + j;
                src[j] = nullptr;
                if (view < T) {
                    const int r = view / V, v = view - r * V;
                    const char* blk = vs.base[r];
                    const size_t wi = (size_t)v * L.words + (n >> 5);
                    const uint32_t w = reinterpret_cast<const uint32_t*>(blk + L.mask_off)[wi];
                    if ((w >> (n & 31)) & 1u) {
                        bits |= 1u << j;
                        const uint32_t row = reinterpret_cast<const uint32_t*>(blk + L.prefix_off)[wi] +
                                             __popc(w & ((1u << (n & 31)) - 1u));
                        src[j] = reinterpret_cast<const float*>(blk + L.rgb_off) + (size_t)row * 3;
                    }
                }
            }
            if (!bits) continue;
            float g[VC][3];
#pragma unroll
            for (int j = 0; j < VC; ++j) {
                g[j][0] = g[j][1] = g[j][2] = 0.f;
                if (bits & (1u << j)) { g[j][0] = src[j][0]; g[j][1] = src[j][1]; g[j][2] = src[j][2]; }
            }
#pragma unroll
            for (int j = 0; j < VC; ++j) {
                const float vr0 = g[j][0], vr1 = g[j][1], vr2 = g[j][2];
                if (vr0 == 0.f && vr1 == 0.f && vr2 == 0.f) continue;
                const float* cp = campos + (base + j) * 4;
                const float dx = m[0] - cp[0], dy = m[1] - cp[1], dz = m[2] - cp[2];
                const float inorm = rsqrt_f(dx * dx + dy * dy + dz * dz);
                float B[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) B[k] = 0.f;
                sh_basis(DEG, dx * inorm, dy * inorm, dz * inorm, B);
#pragma unroll
                for (int gq = 0; gq < NG; ++gq) {
                    float a[12];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 t = acc[3 * gq + q];
                        a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w;
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int k = 4 * gq + b;
                        a[3 * b] = fmaf(B[k], vr0, a[3 * b]);
                        a[3 * b + 1] = fmaf(B[k], vr1, a[3 * b + 1]);
                        a[3 * b + 2] = fmaf(B[k], vr2, a[3 * b + 2]);
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) acc[3 * gq + q] = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
                }
            }
        }
    }
    __syncthreads();
    const int gv = sh_row_floats >> 2;
    const int rows = min(VB, N - n0);
    float4* dst = reinterpret_cast<float4*>(v_sh);
    const float4* src = reinterpret_cast<const float4*>(smem);
    for (int qd = threadIdx.x; qd < rows * gv; qd += VB) {
        const int g = qd / gv, j = qd - g * gv;
        dst[(size_t)(n0 + g) * gv + j] = (j < NV3) ? src[g * ROW + j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int DEG>
static int launch_views(const ViewSrc& vs, int world, int V, int N, const float* means, float* v_sh, int sh_row_floats,
                        cudaStream_t st) {
    const size_t smem = (size_t)VB * 13 * 16 + (size_t)world * V * 16;
    FG_CUDA(cudaFuncSetAttribute(sh_bwd_views_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FG_LAUNCH((sh_bwd_views_kernel<DEG>), ceil_div(N, VB), VB, smem, st, vs, world, V, N, means, v_sh, sh_row_floats);
    return FG_OK;
}

static int make_peers(const fg_xchg_peers* in, Peers& p) {
    FG_REQUIRE(in != nullptr, "peers must not be NULL");
    FG_REQUIRE(in->world >= 1 && in->world <= FG_XCHG_MAX_RANKS, "world must be in 1..FG_XCHG_MAX_RANKS");
    FG_REQUIRE(in->rank >= 0 && in->rank < in->world, "rank out of range");
    p.world = in->world; p.rank = in->rank; p.mc = (char*)in->mc;
    for (int r = 0; r < in->world; ++r) {
        FG_REQUIRE(in->buf[r] && in->flags[r], "peer buffer / flag pointers must not be NULL");
        p.buf[r] = (char*)in->buf[r];
        p.flags[r] = (uint32_t*)in->flags[r];
    }
    return FG_OK;
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_xchg_pub_bytes(int V, int N) {
    const PubLayout L = pub_layout(V, N);
    return (L.rgb_off + (long long)V * N * 12 + 255) / 256 * 256;
}

extern "C" int fg_xchg_pub_layout(int V, int N, int64_t* nnz_off, int64_t* mask_off, int64_t* prefix_off, int64_t* rgb_off,
                                  int32_t* words) {
    FG_REQUIRE(nnz_off && mask_off && prefix_off && rgb_off && words, "NULL pointer");
    const PubLayout L = pub_layout(V, N);
    *nnz_off = L.nnz_off; *mask_off = L.mask_off; *prefix_off = L.prefix_off; *rgb_off = L.rgb_off; *words = L.words;
    return FG_OK;
}

extern "C" int fg_xchg_barrier(const fg_xchg_peers* peers, uint32_t epoch, void* stream) {
    Peers p = {};
    if (int e = make_peers(peers, p)) return e;
    FG_LAUNCH(xchg_barrier_kernel, 1, 32, 0, (cudaStream_t)stream, p, epoch);
    return FG_OK;
}

extern "C" int fg_xchg_allreduce_f32(const fg_xchg_peers* peers, int64_t offset_bytes, int64_t n_floats, uint32_t epoch,
                                     int start_barrier, void* stream) {
    Peers p = {};
    if (int e = make_peers(peers, p)) return e;
    FG_REQUIRE(offset_bytes >= 0 && offset_bytes % 16 == 0, "offset_bytes must be a multiple of 16");
    FG_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "n_floats must be a multiple of 4");
    if (n_floats == 0 && !start_barrier) return FG_OK;
    const long long n4 = n_floats / 4;
    // every rank derives the same grid from the same n: the barriers pair block b with block b
    const long long per = (n4 + p.world - 1) / p.world;
    // in the switch: few fat CTAs (see the kernel); peer loads / stores (2-3 ranks) need the whole GPU's load slots to
    // fill one link: two 512-thread CTAs per SM (measured at 2 ranks, 68 MB: 0.114 ms against 0.205 ms with 32 x 1024)
    // Measured at 8 ranks, 68 MB (tools/bench_xchg_allreduce.py, profiles/r2_allreduce_8gpu.txt): alone, 64 / 128 / 256 CTAs of
    // 256 threads take 0.187 / 0.200 / 0.219 ms (NCCL: 0.281 ms); inside the step, beside the SH-row kernels, 64 CTAs of
    // 1024 threads gave the shortest step (1.95 ms against 2.06 ms with 256 quarter-SM CTAs, which slow those kernels down)
    const int threads = p.mc ? XB : 512;
    const int cap = p.mc ? g_xchg_ar_blocks : 2 * num_sms();
    int grid = (int)std::min<long long>(cap, std::max<long long>(1, (per + threads * 4 - 1) / (threads * 4)));
    FG_REQUIRE(1 + 2 * grid <= FG_XCHG_FLAG_BYTES / (XCHG_SLOT_WORDS * 4), "flag area too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (p.mc) FG_LAUNCH((allreduce_kernel<true>), grid, threads, 0, st, p, (long long)offset_bytes, n4, epoch, start_barrier);
    else FG_LAUNCH((allreduce_kernel<false>), grid, threads, 0, st, p, (long long)offset_bytes, n4, epoch, start_barrier);
    return FG_OK;
}

extern "C" int fg_xchg_sh_bwd_views(const fg_xchg_peers* peers, int64_t pub_offset_bytes, int V, int N, int sh_degree,
                                    int sh_bases, const float* means, float* v_sh, void* staging, int64_t staging_stride,
                                    void* stream) {
    Peers p = {};
    if (int e = make_peers(peers, p)) return e;
    FG_REQUIRE(V >= 1 && N >= 0 && p.world * V <= 256, "views per rank must be >= 1 and world*V <= 256");
    FG_REQUIRE(sh_degree >= 0 && sh_degree <= 3 && sh_bases >= (sh_degree + 1) * (sh_degree + 1), "bad sh_degree / sh_bases");
    FG_REQUIRE(sh_bases % 4 == 0 && (uintptr_t)v_sh % 16 == 0, "v_sh rows must be float4-sized and 16-byte aligned");
    FG_REQUIRE(pub_offset_bytes % 256 == 0, "pub_offset_bytes must be a multiple of 256");
    if (N == 0) return FG_OK;
    FG_REQUIRE(means && v_sh, "means / v_sh must not be NULL");
    FG_REQUIRE(p.world == 1 || (staging && staging_stride >= fg_xchg_pub_bytes(V, N) && staging_stride % 16 == 0 &&
                                (uintptr_t)staging % 16 == 0),
               "staging must hold world blocks of fg_xchg_pub_bytes(V, N) bytes");
    cudaStream_t st = (cudaStream_t)stream;
    const PubLayout L = pub_layout(V, N);
    ViewSrc vs = {};
    for (int r = 0; r < p.world; ++r)
        vs.base[r] = (r == p.rank) ? p.buf[r] + pub_offset_bytes : (const char*)staging + (long long)r * staging_stride;
    if (p.world > 1) {
        // a few CTAs per peer (~3.5 MB per peer on the bench scene), leaving the SMs to the kernels running beside it
        const int bpp = g_xchg_pull_blocks;
        FG_LAUNCH(xchg_pull_kernel, dim3(bpp, p.world), 512, 0, st, p, (long long)pub_offset_bytes, L.nnz_off, L.rgb_off,
                  (char*)staging, (long long)staging_stride);
    }
    switch (sh_degree) {
        case 0: return launch_views<0>(vs, p.world, V, N, means, v_sh, sh_bases * 3, st);
        case 1: return launch_views<1>(vs, p.world, V, N, means, v_sh, sh_bases * 3, st);
        case 2: return launch_views<2>(vs, p.world, V, N, means, v_sh, sh_bases * 3, st);
        default: return launch_views<3>(vs, p.world, V, N, means, v_sh, sh_bases * 3, st);
    }
}
