// Deformation network: the MLP FreeGaussian evaluates for every Gaussian right before
// each render call (freegaussian/freegaussian_model.py:832-845, :1054-1114) -- the one dense contraction of
// the training step, so the one place the 5th-generation tensor cores are used.
//
//   fg_mlp_linear     out = epilogue(A . W^T): CTA = 128 rows x BN outputs, persistent over row tiles.
//                     warp 0 / warp 10 = TMA producers of the activation / weight rings (cp.async.bulk.tensor, mbarriers),
//                     warp 1 = tcgen05.mma issuer (kind::tf32, accumulators in TMEM, double buffered),
//                     warps 2-5 = epilogue (tcgen05.ld -> bias / ReLU / mask -> smem transpose -> coalesced stores),
//                     warps 6-9 = operand split (below).
//                     The reference computes in fp32, so every product runs the error-compensated 3xTF32 scheme:
//                     x = hi + lo with hi the nearest tf32 value (an fp32 with 13 zero low bits) and lo = x - hi, and
//                     each k-step issues A_lo.W_hi + A_hi.W_lo + A_hi.W_hi into the same fp32 accumulator (what is
//                     dropped is O(2^-21) relative).  Activations and gradients stay plain fp32 in HBM (4 bytes per
//                     element each way): the split of the A tile happens in shared memory between TMA and MMA;
//                     the weights are split once per step by fg_mlp_pack.
//   fg_mlp_pack       weights -> padded / reordered / transposed hi+lo operand buffers (one launch, segment table)
//   fg_deform_embed   positional embedding of the means (+ a second point set) + broadcast time embedding -> [N,ld]  (utils.py:27-56)
//   fg_deform_embed_bwd  its VJP with respect to the means (the stage-2 control network does not detach them)
//   fg_deform_apply_fwd/bwd   screw axis -> SE(3) -> means, scales, quats (utils.py:137-159, model.py:841-845)
#include <cuda.h>

#include "common.cuh"

namespace fg {
namespace mlp {

constexpr int BM = 128;      // rows per tile = TMEM lanes
constexpr int BK = 32;       // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;    // tf32 elements per tcgen05.mma
constexpr int A_TILE_BYTES = BM * BK * 4;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a pipeline bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > (1LL << 37)) __trap();  // ~70 s of SM clocks: far beyond any legitimate wait, even time-sliced
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, both operands K-major, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes, stored as TMA's 128-byte swizzle
// writes them: 8-row groups 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
// [cute/arch/mma_sm100_desc.hpp SmemDescriptor: start >> 4 at bit 0, LBO >> 4 at bit 16 (unused for swizzled K-major,
// set to 1), SBO >> 4 at bit 32, version at bit 46, layout type at bit 61]
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor for kind::tf32: D fp32 (bit 4), A and B tf32 (2 at bits 7 and 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.  [cute/arch/mma_sm100_desc.hpp InstrDescriptor]
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// nearest tf32 value (10 mantissa bits) as an fp32 bit pattern with the 13 low bits zero: what the tensor core reads is
// then exact whatever it does with low bits, and x - hi is exact in fp32
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r & 0xFFFFE000u);
}

// ------------------------------------------------------------------------------------------------ the linear layer
struct LinearArgs {
    const float* bias;         // [BN] or NULL
    const uint32_t* mask_in;   // [M, BN/32] bit j of word c = (input of the ReLU at column 32 c + j was > 0)   (EPI_MASK)
    float* out;                // [M, BN]
    uint32_t* mask_out;        // [M, BN/32] (EPI_RELU)
    long long M;
    int kb0, kb1;  // k-blocks read from A0, then from A1 (the skip connection: [h | embedding])
};

enum { EPI_RELU = 0, EPI_LINEAR = 1, EPI_MASK = 2 };

constexpr int kThreads = 320;        // weight-gradient kernel: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-9 operand split
constexpr int STG_PITCH = 36;        // floats per staged row (32 + 4: conflict-free 128-bit writes by row and reads by 4 rows)
constexpr int STG_BYTES = 4 * 32 * STG_PITCH * 4;

constexpr int kLinThreads = 352;     // linear kernel: the same roles + warp 10, the weight-tile TMA producer
constexpr int WK = 16;               // columns per weight tile (64-byte rows, SWIZZLE_64B): two tcgen05.mma k-steps

// K-major tile with 64-byte rows as TMA's 64-byte swizzle writes it: 8-row groups 512 bytes apart, layout type 4
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// Two independent rings, one per source, so that neither stream waits for the other's slots:
//   activation ring: NA slots of [A -> A_hi | A_lo], 128 rows x 32 columns each (from HBM; filled by warp 0, split by warps 6-9)
//   weight ring:     NW slots of [W_hi | W_lo], BN rows x 16 columns each (L2 resident; filled by warp 10)
// (3, 3) is the best split of the 227 KB that was measured -- (2, 4) is 7 % slower -- and equals a single ring of two 96 KB
// stages: at 0.74 of the tf32 MMA peak the tensor pipe, not the loads, is what is left.
template <int BN, int NA, int NW, int EPI>
__global__ void __launch_bounds__(kLinThreads, 1)
    mlp_linear_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl, LinearArgs args) {
    pdl_wait();
    constexpr int A_SLOT_BYTES = 2 * A_TILE_BYTES;
    constexpr int W_TILE_BYTES = BN * WK * 4;
    constexpr int W_SLOT_BYTES = 2 * W_TILE_BYTES;
    constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulators; power of two >= 32
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns");
    static_assert(BN % 32 == 0 && BN <= 256, "BN");
    static_assert(W_SLOT_BYTES % 1024 == 0, "weight slots keep the 1024-byte alignment");
    constexpr int NCHUNK = BN / 32;

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_afull[NA], bar_aconv[NA], bar_aempty[NA], bar_wfull[NW], bar_wempty[NW], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms need 1024-byte alignment
    const uint32_t smemW = smem0 + NA * A_SLOT_BYTES;
    const uint32_t stg0 = smemW + NW * W_SLOT_BYTES;  // epilogue staging, one 32 x 36 tile per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_total = args.kb0 + args.kb1;
    const long long n_tiles = (args.M + BM - 1) / BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) {
            mbar_init(smem_u32(&bar_afull[s]), 1);
            mbar_init(smem_u32(&bar_aconv[s]), 4);
            mbar_init(smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < NW; ++s) {
            mbar_init(smem_u32(&bar_wfull[s]), 1);
            mbar_init(smem_u32(&bar_wempty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&bar_tfull[a]), 1);
            mbar_init(smem_u32(&bar_tempty[a]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== activation TMA producer =====
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int row0 = (int)(tile * BM);
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(smem_u32(&bar_aempty[slot]), phase ^ 1);
                    const uint32_t full = smem_u32(&bar_afull[slot]);
                    mbar_expect_tx(full, A_TILE_BYTES);
                    const bool first = kb < args.kb0;
                    tma_load_2d(smem0 + slot * A_SLOT_BYTES, first ? &mapA0 : &mapA1, full, (first ? kb : kb - args.kb0) * BK, row0);
                    if (++slot == NA) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 10) {
        // ===== weight TMA producer (pre-split hi / lo tiles, L2 resident) =====
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kw = 0; kw < kb_total * (BK / WK); ++kw) {
                    mbar_wait(smem_u32(&bar_wempty[slot]), phase ^ 1);
                    const uint32_t full = smem_u32(&bar_wfull[slot]);
                    const uint32_t sW = smemW + slot * W_SLOT_BYTES;
                    mbar_expect_tx(full, W_SLOT_BYTES);
                    tma_load_2d(sW, &mapWh, full, kw * WK, 0);
                    tma_load_2d(sW + W_TILE_BYTES, &mapWl, full, kw * WK, 0);
                    if (++slot == NW) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): A_lo.W_hi + A_hi.W_lo + A_hi.W_hi per k-step =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int aslot = 0, wslot = 0;
            uint32_t aphase = 0, wphase = 0;
            uint32_t t = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
                const uint32_t acc = t & 1;
                mbar_wait(smem_u32(&bar_tempty[acc]), ((t >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(smem_u32(&bar_aconv[aslot]), aphase);
                    const uint32_t sA = smem0 + aslot * A_SLOT_BYTES;
                    const uint64_t dAh = make_desc(sA), dAl = make_desc(sA + A_TILE_BYTES);
#pragma unroll
                    for (int half = 0; half < BK / WK; ++half) {
                        mbar_wait(smem_u32(&bar_wfull[wslot]), wphase);
                        tc_fence_after();
                        const uint32_t sW = smemW + wslot * W_SLOT_BYTES;
                        const uint64_t dWh = make_desc_sw64(sW), dWl = make_desc_sw64(sW + W_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < WK / UMMA_K; ++k) {
                            // byte offsets inside the swizzled rows, >> 4
                            const uint64_t adv_a = (uint64_t)(((half * (WK / UMMA_K) + k) * UMMA_K * 4) >> 4);
                            const uint64_t adv_w = (uint64_t)((k * UMMA_K * 4) >> 4);
                            tc_mma_tf32(d_tmem, dAl + adv_a, dWh + adv_w, idesc, (kb | half | k) != 0);  // small terms first
                            tc_mma_tf32(d_tmem, dAh + adv_a, dWl + adv_w, idesc, 1);
                            tc_mma_tf32(d_tmem, dAh + adv_a, dWh + adv_w, idesc, 1);
                        }
                        tc_commit(smem_u32(&bar_wempty[wslot]));  // frees the weight slot when these MMAs retire
                        if (++wslot == NW) { wslot = 0; wphase ^= 1; }
                    }
                    tc_commit(smem_u32(&bar_aempty[aslot]));
                    if (++aslot == NA) { aslot = 0; aphase ^= 1; }
                }
                tc_commit(smem_u32(&bar_tfull[acc]));  // accumulator complete
            }
        }
    } else if (warp < 6) {
        // ===== epilogue: warp w may touch TMEM lanes 32 (w % 4) .. +31; lane = row =====
        const int q = warp & 3;
        const uint32_t stg = stg0 + q * (32 * STG_PITCH * 4);
        uint32_t t = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const uint32_t acc = t & 1;
            const long long wrow0 = tile * BM + q * 32;  // first row of this warp
            uint32_t bits[NCHUNK];
            if (EPI == EPI_MASK) {
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) bits[c] = (wrow0 + lane < args.M) ? __ldg(args.mask_in + (wrow0 + lane) * NCHUNK + c) : 0u;
            }
            mbar_wait(smem_u32(&bar_tfull[acc]), (t >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, v);
                if (EPI == EPI_RELU) {
                    uint32_t m = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = __uint_as_float(v[j]) + __ldg(args.bias + c * 32 + j);
                        m |= (x > 0.f ? 1u : 0u) << j;
                        v[j] = __float_as_uint(fmaxf(x, 0.f));
                    }
                    bits[c] = m;
                } else if (EPI == EPI_LINEAR) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(args.bias + c * 32 + j));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (bits[c] >> j) & 1u ? v[j] : 0u;
                }
                // transpose through shared memory: written by row (lane = row), read back four rows per instruction so that
                // every global store instruction covers four complete 128-byte lines
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (lane * STG_PITCH + 4 * j) * 4), "r"(v[4 * j]),
                                 "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                 : "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3), cc = (lane & 7) * 4;
                    uint4 o;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                                 : "r"(stg + (rr * STG_PITCH + cc) * 4)
                                 : "memory");
                    if (wrow0 + rr < args.M) *reinterpret_cast<uint4*>(args.out + (wrow0 + rr) * BN + c * 32 + cc) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
            if constexpr (EPI == EPI_RELU) {
                if (wrow0 + lane < args.M)
#pragma unroll
                for (int c = 0; c < NCHUNK; c += 4)
                    *reinterpret_cast<uint4*>(args.mask_out + (wrow0 + lane) * NCHUNK + c) = make_uint4(bits[c], bits[c + 1], bits[c + 2], bits[c + 3]);
            }
        }
    } else if (warp < 10) {
        // ===== operand split: the landed fp32 tile becomes hi (in place) and lo (next tile); element-wise, so the swizzle
        // TMA applied is irrelevant -- lo lands at the same swizzled offset of its own tile =====
        const int tid = threadIdx.x - 192;
        int slot = 0;
        uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < kb_total; ++kb) {
                mbar_wait(smem_u32(&bar_afull[slot]), phase);
                const uint32_t sA = smem0 + slot * A_SLOT_BYTES;
#pragma unroll
                for (int i = 0; i < A_TILE_BYTES / 16 / 128; ++i) {
                    const uint32_t addr = sA + (i * 128 + tid) * 16;
                    float4 x;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr) : "memory");
                    const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr + A_TILE_BYTES), "f"(x.x - h.x), "f"(x.y - h.y),
                                 "f"(x.z - h.z), "f"(x.w - h.w)
                                 : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_aconv[slot]));
                if (++slot == NA) { slot = 0; phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows, cols] fp32 row-major (pitch = cols), box = box_rows x box_cols columns; rows past the end read 0
static bool make_map(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows,
                     CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int box_cols = BK) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int NA, int NW, int EPI>
static int launch_linear(long long M, const float* a0, int k0, const float* a1, int k1, const float* wh, const float* wl,
                         const LinearArgs& args, cudaStream_t st) {
    CUtensorMap mA0, mA1, mWh, mWl;
    bool ok = make_map(&mA0, a0, M, k0, BM) && make_map(&mWh, wh, BN, k0 + k1, BN, CU_TENSOR_MAP_SWIZZLE_64B, WK) &&
              make_map(&mWl, wl, BN, k0 + k1, BN, CU_TENSOR_MAP_SWIZZLE_64B, WK);
    mA1 = mA0;
    if (k1 > 0) ok = ok && make_map(&mA1, a1, M, k1, BM);
    if (!ok) return set_error(FG_ERR_CUDA, "cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)", __FILE__, __LINE__);
    constexpr int SMEM = NA * 2 * A_TILE_BYTES + NW * 2 * BN * WK * 4 + STG_BYTES + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    auto kern = mlp_linear_kernel<BN, NA, NW, EPI>;
    FG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const long long n_tiles = (M + BM - 1) / BM;
    const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
    FG_LAUNCH(kern, grid, kLinThreads, SMEM, st, mA0, mA1, mWh, mWl, args);
    return FG_OK;
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[256, KIN] += dz^T . a   and   db[256] += column sums of dz,   dz [N, 256], a [N, KIN], both row-major fp32.
// The contraction runs over the rows, so both operands are "MN-major" for the tensor core: a TMA box of 16 rows x 32
// columns lands as 16 swizzled 128-byte rows = four 4-row atoms of the canonical MN-major layout (make_desc_mn).  Split-K: CTA c owns a contiguous range of row blocks, accumulates the whole
// [256, KIN] product in TMEM (two 128-lane halves x KIN columns) and adds it to dW with vector reductions at the end.
// Same 3xTF32 scheme and on-chip operand split as the linear layer; the split warps also accumulate db.
constexpr int WG_ROWS = 16;                       // rows (K extent) per stage = two tcgen05.mma k-steps
constexpr int WG_CHUNK_BYTES = WG_ROWS * 128;     // one 32-column chunk of a stage
constexpr int WG_DZ_BYTES = 8 * WG_CHUNK_BYTES;   // [16, 256] fp32

struct WgradArgs {
    float* dw;       // [256, ld_dw], written at column col0
    float* db;       // [256] or NULL
    long long N;
    int ld_dw, col0;
};

__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) { return make_idesc(m, n) | (1u << 15) | (1u << 16); }

// 32-bit MN-major operands have exactly one legal shared-memory layout, SWIZZLE_128B_BASE32B (layout type 1;
// cutlass/gemm/collective/builders/sm100_common.inl "for mn-major tf32 operands, SW128_32B is the only available smem
// layout"): rows of 128 bytes (32 columns), the 32-byte chunk index XORed with (row % 4) -- what TMA's
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes -- atoms of 4 rows (SBO = 512 bytes between them), LBO = distance between
// 32-column atoms.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(WG_CHUNK_BYTES >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
           (1ull << 61);
}

__device__ __forceinline__ void red_add_v4(float* p, uint4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

template <int KIN, int NSTAGE>
__global__ void __launch_bounds__(kThreads, 1)
    mlp_wgrad_kernel(const __grid_constant__ CUtensorMap mapDz, const __grid_constant__ CUtensorMap mapA, WgradArgs args) {
    pdl_wait();
    constexpr int A_BYTES = (KIN / 32) * WG_CHUNK_BYTES;
    constexpr int STAGE_BYTES = 2 * (WG_DZ_BYTES + A_BYTES);  // [dz -> dz_hi | dz_lo | a -> a_hi | a_lo]
    constexpr int TMEM_COLS = 2 * KIN <= 256 ? 256 : 512;
    static_assert(KIN % 32 == 0 && KIN <= 256, "KIN");
    constexpr int NCHUNK = KIN / 32;

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[NSTAGE], bar_conv[NSTAGE], bar_empty[NSTAGE], bar_done;
    __shared__ uint32_t tmem_base_slot;

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stg0 = smem0 + NSTAGE * STAGE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // contiguous range of 16-row blocks of this CTA
    const long long n_blocks = (args.N + WG_ROWS - 1) / WG_ROWS;
    const long long per = (n_blocks + gridDim.x - 1) / gridDim.x;
    const long long blk0 = (long long)blockIdx.x * per;
    const long long blk1 = blk0 + per < n_blocks ? blk0 + per : n_blocks;
    const int n_kb = blk1 > blk0 ? (int)(blk1 - blk0) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_conv[s]), 4);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (n_kb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
                    const uint32_t full = smem_u32(&bar_full[stage]);
                    const uint32_t sD = smem0 + stage * STAGE_BYTES;
                    const uint32_t sA = sD + 2 * WG_DZ_BYTES;
                    const int row0 = (int)((blk0 + kb) * WG_ROWS);
                    mbar_expect_tx(full, WG_DZ_BYTES + A_BYTES);
#pragma unroll
                    for (int c = 0; c < 8; ++c) tma_load_2d(sD + c * WG_CHUNK_BYTES, &mapDz, full, c * 32, row0);
#pragma unroll
                    for (int c = 0; c < NCHUNK; ++c) tma_load_2d(sA + c * WG_CHUNK_BYTES, &mapA, full, c * 32, row0);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc_mn(128, KIN);
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(smem_u32(&bar_conv[stage]), phase);
                    tc_fence_after();
                    const uint32_t sD = smem0 + stage * STAGE_BYTES;
                    const uint32_t sA = sD + 2 * WG_DZ_BYTES;
#pragma unroll
                    for (int k = 0; k < WG_ROWS / UMMA_K; ++k) {
                        const uint64_t bh = make_desc_mn(sA + k * 1024), bl = make_desc_mn(sA + A_BYTES + k * 1024);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {  // output rows 128 h .. 128 h + 127 = dz columns = chunks 4 h .. 4 h + 3
                            const uint64_t ah = make_desc_mn(sD + h * 4 * WG_CHUNK_BYTES + k * 1024);
                            const uint64_t al = make_desc_mn(sD + WG_DZ_BYTES + h * 4 * WG_CHUNK_BYTES + k * 1024);
                            const uint32_t d = tmem_base + h * KIN;
                            tc_mma_tf32(d, al, bh, idesc, (kb | k) != 0);
                            tc_mma_tf32(d, ah, bl, idesc, 1);
                            tc_mma_tf32(d, ah, bh, idesc, 1);
                        }
                    }
                    tc_commit(smem_u32(&bar_empty[stage]));
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc_commit(smem_u32(&bar_done));
            }
        } else if (warp < 6) {
            // ===== epilogue: TMEM lane = output row (dz column), TMEM column = input feature =====
            const int q = warp & 3;
            const uint32_t stg = stg0 + q * (32 * STG_PITCH * 4);
            mbar_wait(smem_u32(&bar_done), 0);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int orow0 = h * 128 + q * 32;
#pragma unroll 1
                for (int c = 0; c < NCHUNK; ++c) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * KIN + c * 32, v);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (lane * STG_PITCH + 4 * j) * 4), "r"(v[4 * j]),
                                     "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                     : "memory");
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + (lane >> 3), cc = (lane & 7) * 4;
                        uint4 o;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                                     : "r"(stg + (rr * STG_PITCH + cc) * 4)
                                     : "memory");
                        red_add_v4(args.dw + (long long)(orow0 + rr) * args.ld_dw + args.col0 + c * 32 + cc, o);
                    }
                }
            }
            tc_fence_before();
        } else {
            // ===== operand split of both tiles (+ column sums of dz for the bias gradient) =====
            const int tid = threadIdx.x - 192;
            // this thread always sees row r = tid / 8 of a stage and the 16-byte slot s = tid % 8 of every chunk; the 32-byte
            // chunk s / 2 holds logical chunk (s / 2) ^ (r % 4), so the thread owns columns 32 c + col4 .. + 3 of chunk c
            const int col4 = 4 * (((((tid & 7) >> 1) ^ ((tid >> 3) & 3)) << 1) | (tid & 1));
            float4 sum[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) sum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(smem_u32(&bar_full[stage]), phase);
                const uint32_t sD = smem0 + stage * STAGE_BYTES;
                const uint32_t sA = sD + 2 * WG_DZ_BYTES;
#pragma unroll
                for (int c = 0; c < 8 + NCHUNK; ++c) {
                    const bool is_dz = c < 8;
                    const uint32_t addr = (is_dz ? sD + c * WG_CHUNK_BYTES : sA + (c - 8) * WG_CHUNK_BYTES) + tid * 16;
                    float4 x;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr) : "memory");
                    if (is_dz) { sum[c & 7].x += x.x; sum[c & 7].y += x.y; sum[c & 7].z += x.z; sum[c & 7].w += x.w; }
                    const float4 hh = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(hh.x), "f"(hh.y), "f"(hh.z), "f"(hh.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr + (is_dz ? WG_DZ_BYTES : A_BYTES)), "f"(x.x - hh.x),
                                 "f"(x.y - hh.y), "f"(x.z - hh.z), "f"(x.w - hh.w)
                                 : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_conv[stage]));
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (args.db) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    // reduce over the 4 rows of this warp (lane = 8 (r % 4) + s) that hold the same columns, then one atomic
                    // per column
                    float4 t = sum[c];
#pragma unroll
                    for (int ofs = 8; ofs < 32; ofs <<= 1) {
                        // partner row r' = r ^ (ofs / 8) keeps the same columns in slot s' = s ^ 2 (ofs / 8)
                        const int src = lane ^ ofs ^ (ofs >> 2);
                        t.x += __shfl_sync(0xffffffffu, t.x, src);
                        t.y += __shfl_sync(0xffffffffu, t.y, src);
                        t.z += __shfl_sync(0xffffffffu, t.z, src);
                        t.w += __shfl_sync(0xffffffffu, t.w, src);
                    }
                    if ((lane >> 3) == 0) {
                        float* p = args.db + c * 32 + col4;
                        atomicAdd(p, t.x); atomicAdd(p + 1, t.y); atomicAdd(p + 2, t.z); atomicAdd(p + 3, t.w);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

template <int KIN, int NSTAGE>
static int launch_wgrad(long long N, const float* dz, const float* a, const WgradArgs& args, cudaStream_t st) {
    CUtensorMap mDz, mA;
    if (!(make_map(&mDz, dz, N, 256, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
          make_map(&mA, a, N, KIN, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)))
        return set_error(FG_ERR_CUDA, "cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)", __FILE__, __LINE__);
    constexpr int STAGE_BYTES = 2 * (WG_DZ_BYTES + (KIN / 32) * WG_CHUNK_BYTES);
    constexpr int SMEM = NSTAGE * STAGE_BYTES + STG_BYTES + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    auto kern = mlp_wgrad_kernel<KIN, NSTAGE>;
    FG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const long long n_blocks = (N + WG_ROWS - 1) / WG_ROWS;
    const long long want = (n_blocks + 63) / 64;  // at least 64 row blocks (1024 rows) per CTA
    const int grid = (int)(want < 1 ? 1 : (want < num_sms() ? want : num_sms()));
    FG_LAUNCH(kern, grid, kThreads, SMEM, st, mDz, mA, args);
    return FG_OK;
}

// ------------------------------------------------------------------------------------------------ weight packing
struct PackTable {
    fg_mlp_pack_segment seg[FG_MLP_PACK_MAX_SEGMENTS];
};

__global__ void __launch_bounds__(256) mlp_pack_kernel(PackTable tab) {
    pdl_wait();
    const fg_mlp_pack_segment s = tab.seg[blockIdx.y];
    const int total = s.rows * s.cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        // i runs over the destination so that the stores coalesce
        int r, c;
        float x;
        if (s.transpose) {  // dst[c_src, r_src]
            r = i / s.rows;  // source column
            c = i % s.rows;  // source row
            x = s.src[(long long)c * s.src_ld + s.src_col0 + r];
        } else {
            r = i / s.cols;
            c = i % s.cols;
            x = s.src[(long long)r * s.src_ld + s.src_col0 + c];
        }
        const long long d = (long long)r * s.dst_ld + s.dst_col0 + c;
        const float h = tf32_hi(x);
        s.dst_hi[d] = h;
        if (s.dst_lo) s.dst_lo[d] = x - h;
    }
}

// ------------------------------------------------------------------------------------------------ embedding
// E[n, :] = [x, sin(x 2^0), cos(x 2^0), ..., sin(x 2^9), cos(x 2^9) | t_emb | 0...]   (utils.py:27-56; model.py:1095-1096)
// thread = (row, unit): units 0 .. multires-1 of a point set are its frequencies (one sincosf per coordinate -> the six
// columns [sin | cos] of that frequency), the following unit copies the raw coordinates; after the point sets come the
// t_emb / zero-padding columns, one unit per column.  Row layout: embed(x) | embed(x2) (if given) | t_emb | 0.
__global__ void __launch_bounds__(256) deform_embed_kernel(long long N, const float* __restrict__ x, const float* __restrict__ x2,
                                                           const float* __restrict__ t_emb, int t_ch, int multires, int ld,
                                                           float* __restrict__ e) {
    pdl_wait();
    const int x_ch = 3 + 6 * multires;
    const int sets = x2 ? 2 : 1;
    const int p_ch = sets * x_ch;
    const int units = sets * (multires + 1) + (ld - p_ch);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * units) return;
    const long long n = i / units;
    int u = (int)(i % units);
    float* row = e + n * ld;
    if (u < sets * (multires + 1)) {
        const int set = u / (multires + 1), f = u % (multires + 1);
        const float* src = (set ? x2 : x) + n * 3;
        float* dst = row + set * x_ch;
        if (f == multires) {
            dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
        } else {
            const float w = exp2f((float)f);  // freq = 2^f exactly, so x * freq is the reference's product
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float sn, cs;
                sincosf(src[c] * w, &sn, &cs);
                dst[3 + 6 * f + c] = sn;
                dst[3 + 6 * f + 3 + c] = cs;
            }
        }
    } else {
        const int j = p_ch + (u - sets * (multires + 1));
        row[j] = j < p_ch + t_ch ? t_emb[j - p_ch] : 0.f;
    }
}

// VJP of embed(x) (the first 3 + 6 multires columns of a row): dx = de_x + sum_f 2^f (cos(x 2^f) de_sin,f - sin(x 2^f) de_cos,f)
__global__ void __launch_bounds__(256) deform_embed_bwd_kernel(long long N, const float* __restrict__ x, const float* __restrict__ de,
                                                               int multires, int ld, float* __restrict__ dx) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 3) return;
    const long long n = i / 3;
    const int c = (int)(i % 3);
    const float* row = de + n * ld;
    const float xv = x[i];
    float g = row[c];
    for (int f = 0; f < multires; ++f) {
        const float w = exp2f((float)f);
        float sn, cs;
        sincosf(xv * w, &sn, &cs);
        g += w * (cs * row[3 + 6 * f + c] - sn * row[3 + 6 * f + 3 + c]);
    }
    dx[i] = g;
}

// ------------------------------------------------------------------------------------------------ SE(3) application
struct Screw {
    float th, w[3], v[3], sn, cs, W[9], W2[9], R[9], G[9];
};

__device__ __forceinline__ void screw_from_head(const float* o, Screw& s) {
    // model.py:1103-1109: theta = |w|, w = w / theta + 1e-5, v = v / theta + 1e-5; utils.py:137-159
    s.th = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s.w[k] = o[k] / s.th + 1e-5f;
        s.v[k] = o[3 + k] / s.th + 1e-5f;
    }
    sincosf(s.th, &s.sn, &s.cs);
    const float w0 = s.w[0], w1 = s.w[1], w2 = s.w[2];
    const float W[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s.W[r * 3 + c] = W[r * 3 + c];
            s.W2[r * 3 + c] = W[r * 3 + 0] * W[0 * 3 + c] + W[r * 3 + 1] * W[1 * 3 + c] + W[r * 3 + 2] * W[2 * 3 + c];
        }
    const float c1 = 1.f - s.cs, c2 = s.th - s.sn;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float eye = (k % 4 == 0) ? 1.f : 0.f;
        s.R[k] = eye + s.sn * s.W[k] + c1 * s.W2[k];
        s.G[k] = s.th * eye + c1 * s.W[k] + c2 * s.W2[k];
    }
}

__global__ void __launch_bounds__(256)
    deform_apply_fwd_kernel(long long N, const float* __restrict__ head, const float* __restrict__ means, const float* __restrict__ scales_log,
                            const float* __restrict__ quats, float* __restrict__ means_out, float* __restrict__ scales_out,
                            float* __restrict__ quats_out) {
    pdl_wait();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float o[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(head + n * FG_MLP_HEAD_LD) + k);
        o[4 * k] = t.x, o[4 * k + 1] = t.y, o[4 * k + 2] = t.z, o[4 * k + 3] = t.w;
    }
    Screw s;
    screw_from_head(o, s);
    const float m0 = means[n * 3], m1 = means[n * 3 + 1], m2 = means[n * 3 + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float p = s.G[r * 3] * s.v[0] + s.G[r * 3 + 1] * s.v[1] + s.G[r * 3 + 2] * s.v[2];
        means_out[n * 3 + r] = s.R[r * 3] * m0 + s.R[r * 3 + 1] * m1 + s.R[r * 3 + 2] * m2 + p;  // model.py:841 (w row = 0,0,0,1)
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) scales_out[n * 3 + k] = expf(scales_log[n * 3 + k]) + o[10 + k];  // model.py:844
    const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + n);
    const float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float4 qo;
    qo.x = q.x / qn + o[6], qo.y = q.y / qn + o[7], qo.z = q.z / qn + o[8], qo.w = q.w / qn + o[9];  // model.py:845
    reinterpret_cast<float4*>(quats_out)[n] = qo;
}

__global__ void __launch_bounds__(256)
    deform_apply_bwd_kernel(long long N, const float* __restrict__ head, const float* __restrict__ means, const float* __restrict__ scales_log,
                            const float* __restrict__ quats, const float* __restrict__ v_means_out, const float* __restrict__ v_scales_out,
                            const float* __restrict__ v_quats_out, float* __restrict__ v_head, float* __restrict__ v_means,
                            float* __restrict__ v_scales_log, float* __restrict__ v_quats) {
    pdl_wait();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float o[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(head + n * FG_MLP_HEAD_LD) + k);
        o[4 * k] = t.x, o[4 * k + 1] = t.y, o[4 * k + 2] = t.z, o[4 * k + 3] = t.w;
    }
    Screw s;
    screw_from_head(o, s);
    const float m[3] = {means[n * 3], means[n * 3 + 1], means[n * 3 + 2]};
    const float g[3] = {v_means_out[n * 3], v_means_out[n * 3 + 1], v_means_out[n * 3 + 2]};
    // means' = R m + G v
#pragma unroll
    for (int c = 0; c < 3; ++c) v_means[n * 3 + c] = s.R[c] * g[0] + s.R[3 + c] * g[1] + s.R[6 + c] * g[2];
    float gR[9], gG[9], gv[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gR[r * 3 + c] = g[r] * m[c];
            gG[r * 3 + c] = g[r] * s.v[c];
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) gv[c] = s.G[c] * g[0] + s.G[3 + c] * g[1] + s.G[6 + c] * g[2];
    const float c1 = 1.f - s.cs, c2 = s.th - s.sn;
    // dL/dW = sn gR + c1 gG + c1 (gR W^T + W^T gR) + c2 (gG W^T + W^T gG);   d(W W) pulls back as X W^T + W^T X
    float gW[9];
    float dot_RW = 0.f, dot_RW2 = 0.f, dot_GW = 0.f, dot_GW2 = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float xr = 0.f, xg = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                xr += gR[r * 3 + k] * s.W[c * 3 + k] + s.W[k * 3 + r] * gR[k * 3 + c];
                xg += gG[r * 3 + k] * s.W[c * 3 + k] + s.W[k * 3 + r] * gG[k * 3 + c];
            }
            gW[r * 3 + c] = s.sn * gR[r * 3 + c] + c1 * gG[r * 3 + c] + c1 * xr + c2 * xg;
            dot_RW += gR[r * 3 + c] * s.W[r * 3 + c];
            dot_RW2 += gR[r * 3 + c] * s.W2[r * 3 + c];
            dot_GW += gG[r * 3 + c] * s.W[r * 3 + c];
            dot_GW2 += gG[r * 3 + c] * s.W2[r * 3 + c];
        }
    const float g_th = s.cs * dot_RW + s.sn * dot_RW2 + (gG[0] + gG[4] + gG[8]) + s.sn * dot_GW + c1 * dot_GW2;
    const float gw[3] = {gW[7] - gW[5], gW[2] - gW[6], gW[3] - gW[1]};
    const float inv = 1.f / s.th;
    const float g_th_total = g_th - (gw[0] * o[0] + gw[1] * o[1] + gw[2] * o[2]) * inv * inv -
                             (gv[0] * o[3] + gv[1] * o[4] + gv[2] * o[5]) * inv * inv;
    float vo[FG_MLP_HEAD_LD];
#pragma unroll
    for (int k = 0; k < FG_MLP_HEAD_LD; ++k) vo[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vo[k] = gw[k] * inv + g_th_total * o[k] * inv;
        vo[3 + k] = gv[k] * inv;
        const float gs = v_scales_out[n * 3 + k];
        vo[10 + k] = gs;
        v_scales_log[n * 3 + k] = gs * expf(scales_log[n * 3 + k]);
    }
    const float4 q = __ldg(reinterpret_cast<const float4*>(quats) + n);
    const float4 gq = __ldg(reinterpret_cast<const float4*>(v_quats_out) + n);
    vo[6] = gq.x, vo[7] = gq.y, vo[8] = gq.z, vo[9] = gq.w;
    const float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float qh[4] = {q.x / qn, q.y / qn, q.z / qn, q.w / qn};
    const float d = qh[0] * gq.x + qh[1] * gq.y + qh[2] * gq.z + qh[3] * gq.w;
    float4 vq;
    vq.x = (gq.x - qh[0] * d) / qn, vq.y = (gq.y - qh[1] * d) / qn, vq.z = (gq.z - qh[2] * d) / qn, vq.w = (gq.w - qh[3] * d) / qn;
    reinterpret_cast<float4*>(v_quats)[n] = vq;
#pragma unroll
    for (int k = 0; k < FG_MLP_HEAD_LD / 4; ++k)
        reinterpret_cast<float4*>(v_head + n * FG_MLP_HEAD_LD)[k] = make_float4(vo[4 * k], vo[4 * k + 1], vo[4 * k + 2], vo[4 * k + 3]);
}

}  // namespace mlp
}  // namespace fg

// ------------------------------------------------------------------------------------------------ C ABI
using namespace fg;
using namespace fg::mlp;

extern "C" int fg_mlp_linear(int mode, int64_t M, int n_out, const float* a0, int k0, const float* a1, int k1, const float* w_hi,
                             const float* w_lo, const float* bias, const uint32_t* mask_in, float* out, uint32_t* mask_out,
                             void* stream) {
    FG_REQUIRE(M >= 0 && M < (1ll << 31) - BM, "fg_mlp_linear: M out of range");
    FG_REQUIRE(k0 > 0 && k0 % BK == 0 && k1 >= 0 && k1 % BK == 0, "fg_mlp_linear: k0, k1 must be multiples of 32 (k0 > 0)");
    if (M == 0) return FG_OK;  // empty inputs have NULL data pointers
    FG_REQUIRE(a0 && w_hi && w_lo && out && (k1 == 0 || a1), "fg_mlp_linear: NULL operand");
    LinearArgs args = {bias, mask_in, out, mask_out, (long long)M, k0 / BK, k1 / BK};
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == FG_MLP_RELU) {
        FG_REQUIRE(n_out == 256 && bias && mask_out, "fg_mlp_linear: FG_MLP_RELU is built for 256 outputs and needs bias and mask_out");
        return launch_linear<256, 3, 3, EPI_RELU>(M, a0, k0, a1, k1, w_hi, w_lo, args, st);
    }
    if (mode == FG_MLP_LINEAR) {
        FG_REQUIRE((n_out == FG_MLP_HEAD_LD || n_out == 128) && bias, "fg_mlp_linear: FG_MLP_LINEAR is built for 32 or 128 outputs and needs a bias");
        if (n_out == 128) return launch_linear<128, 3, 3, EPI_LINEAR>(M, a0, k0, a1, k1, w_hi, w_lo, args, st);
        return launch_linear<FG_MLP_HEAD_LD, 4, 4, EPI_LINEAR>(M, a0, k0, a1, k1, w_hi, w_lo, args, st);
    }
    if (mode == FG_MLP_DGRAD) {
        FG_REQUIRE(n_out == 256 && mask_in, "fg_mlp_linear: FG_MLP_DGRAD is built for 256 outputs and needs mask_in");
        return launch_linear<256, 3, 3, EPI_MASK>(M, a0, k0, a1, k1, w_hi, w_lo, args, st);
    }
    return set_error(FG_ERR_INVALID, "fg_mlp_linear: unknown mode", __FILE__, __LINE__);
}

extern "C" int fg_mlp_wgrad(int64_t N, const float* dz, const float* a, int k_in, float* dw, int ld_dw, int col0, float* db, void* stream) {
    FG_REQUIRE(N >= 0 && N < (1ll << 31) - 256, "fg_mlp_wgrad: N out of range");
    FG_REQUIRE(col0 >= 0 && col0 % 4 == 0 && ld_dw % 4 == 0 && col0 + k_in <= ld_dw, "fg_mlp_wgrad: dw columns must be 16-byte aligned and in range");
    if (N == 0) return FG_OK;  // empty inputs have NULL data pointers
    FG_REQUIRE(dz && a && dw, "fg_mlp_wgrad: NULL operand");
    WgradArgs args = {dw, db, (long long)N, ld_dw, col0};
    cudaStream_t st = (cudaStream_t)stream;
    if (k_in == 256) return launch_wgrad<256, 3>(N, dz, a, args, st);
    if (k_in == FG_MLP_EMBED_LD) return launch_wgrad<FG_MLP_EMBED_LD, 4>(N, dz, a, args, st);
    if (k_in == FG_MLP_HEAD_LD) return launch_wgrad<FG_MLP_HEAD_LD, 4>(N, dz, a, args, st);
    if (k_in == 128) return launch_wgrad<128, 3>(N, dz, a, args, st);
    return set_error(FG_ERR_INVALID, "fg_mlp_wgrad: k_in must be 256, 128, 96 or 32", __FILE__, __LINE__);
}

extern "C" int fg_mlp_pack(int n_segments, const fg_mlp_pack_segment* segments_host, void* stream) {
    FG_REQUIRE(n_segments >= 0 && n_segments <= FG_MLP_PACK_MAX_SEGMENTS, "fg_mlp_pack: too many segments");
    if (n_segments == 0) return FG_OK;
    PackTable tab;
    for (int i = 0; i < n_segments; ++i) {
        const fg_mlp_pack_segment& s = segments_host[i];
        FG_REQUIRE(s.src && s.dst_hi && s.rows > 0 && s.cols > 0, "fg_mlp_pack: bad segment");
        tab.seg[i] = s;
    }
    FG_LAUNCH(mlp_pack_kernel, dim3(64, n_segments), 256, 0, (cudaStream_t)stream, tab);
    return FG_OK;
}

extern "C" int fg_deform_embed(int64_t N, const float* x, const float* x2, const float* t_emb, int t_ch, int multires, int ld, float* e,
                               void* stream) {
    FG_REQUIRE(N >= 0 && multires >= 0 && t_ch >= 0 && ld > 0 && ld % 32 == 0, "fg_deform_embed: bad sizes (ld must be a multiple of 32)");
    FG_REQUIRE((x2 ? 2 : 1) * (3 + 6 * multires) + t_ch <= ld && (t_ch == 0 || t_emb), "fg_deform_embed: embedding wider than ld");
    if (N == 0) return FG_OK;
    FG_REQUIRE(x && e, "fg_deform_embed: NULL argument");
    const int units = (x2 ? 2 : 1) * (multires + 1) + (ld - (x2 ? 2 : 1) * (3 + 6 * multires));
    FG_LAUNCH(deform_embed_kernel, ceil_div(N * units, 256), 256, 0, (cudaStream_t)stream, (long long)N, x, x2, t_emb, t_ch, multires, ld, e);
    return FG_OK;
}

extern "C" int fg_deform_embed_bwd(int64_t N, const float* x, const float* de, int multires, int ld, float* dx, void* stream) {
    FG_REQUIRE(N >= 0 && multires >= 0 && 3 + 6 * multires <= ld, "fg_deform_embed_bwd: bad sizes");
    if (N == 0) return FG_OK;
    FG_REQUIRE(x && de && dx, "fg_deform_embed_bwd: NULL argument");
    FG_LAUNCH(deform_embed_bwd_kernel, ceil_div(N * 3, 256), 256, 0, (cudaStream_t)stream, (long long)N, x, de, multires, ld, dx);
    return FG_OK;
}

extern "C" int fg_deform_apply_fwd(int64_t N, const float* head, const float* means, const float* scales_log, const float* quats,
                                   float* means_out, float* scales_out, float* quats_out, void* stream) {
    FG_REQUIRE(N >= 0 && head && means && scales_log && quats && means_out && scales_out && quats_out, "fg_deform_apply_fwd: NULL argument");
    if (N == 0) return FG_OK;
    FG_LAUNCH(deform_apply_fwd_kernel, ceil_div(N, 256), 256, 0, (cudaStream_t)stream, (long long)N, head, means, scales_log, quats,
              means_out, scales_out, quats_out);
    return FG_OK;
}

extern "C" int fg_deform_apply_bwd(int64_t N, const float* head, const float* means, const float* scales_log, const float* quats,
                                   const float* v_means_out, const float* v_scales_out, const float* v_quats_out, float* v_head,
                                   float* v_means, float* v_scales_log, float* v_quats, void* stream) {
    FG_REQUIRE(N >= 0 && head && means && scales_log && quats && v_means_out && v_scales_out && v_quats_out && v_head && v_means &&
                   v_scales_log && v_quats,
               "fg_deform_apply_bwd: NULL argument");
    if (N == 0) return FG_OK;
    FG_LAUNCH(deform_apply_bwd_kernel, ceil_div(N, 256), 256, 0, (cudaStream_t)stream, (long long)N, head, means, scales_log, quats,
              v_means_out, v_scales_out, v_quats_out, v_head, v_means, v_scales_log, v_quats);
    return FG_OK;
}
