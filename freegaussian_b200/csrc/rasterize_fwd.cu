// (3a) Per-tile front-to-back alpha compositing, forward: RGB + depth + flow (all CH
// channels) in one pass.  Replaces gsplat rasterize_to_pixels_fwd
// behind freegaussian_model.py:847-868.
//
// Roofline: FP32 pipe / shared memory.  Unit = one evaluated (pixel, Gaussian) pair:
// ~12 flop for delta/sigma/alpha/T + 2 flop per channel + one MUFU.EX2 (24 flop at 6
// channels).  HBM floor: 32-56 B gathered per intersection + the images.
//
// One CTA per 16x16 tile, one pixel per thread, warps own 8x4 patches.  The tile's sorted
// list is consumed in batches of 256 Gaussians staged in shared memory as float4 records
// (every inner-loop read is a broadcast LDS.128).
#include <string.h>

#include "rasterize_common.cuh"

namespace fg {

struct RasterFwdParams {
    int C, N, width, height, tile_w, tile_h;
    const float2* means2d;
    const float* conics;
    const float* feat;
    const float* opacities;
    const float* backgrounds;
    const float4* flow_affine;
    int flow_ch0;
    int split;       // channels [0,split) -> render, [split,CH) -> render2
    int ed_channel;  // this channel is divided by max(alpha, 1e-10) ("ED"), -1 = none
    int opac_shared; // opacities is [N] shared by all cameras (index g % N) instead of [C*N]
    const int32_t* isect_offsets;
    const int32_t* flatten_ids;
    long long n_isects;
    float* render;
    float* render2;
    float* alphas;
    int32_t* last_ids;
};

// which forward kernel fg_rasterize_fwd launches (fg_set_option("fwd_two_pixels", 0 / 1)); both give the same images.
// Default: the one-pixel kernel.  Measured on cfg3 (profiles/r2_rasterize_fwd2_kernel.txt): the two-pixel packed kernel
// executes 5 % fewer instructions but its per-pixel predication (validity, stop test, selects of w / T / last id: FSETP,
// FSEL, FMNMX) lands on the half-rate ALU pipe -- 71 % busy, math-pipe-throttle stalls -- and it runs 0.296 ms against
// 0.270 ms.  (The backward kernel has no such per-pixel control flow and gains 8.6 % from the same packing.)
int g_fwd_two_pixels = 0;

template <int CH, bool AFF>
__global__ void __launch_bounds__(TILE_PIX, 6) rasterize_fwd_kernel(RasterFwdParams p) {
    pdl_wait();
    constexpr int FV = (CH + 3) / 4;  // float4 per Gaussian for the features
    constexpr int NREC = 2 + FV + (AFF ? 1 : 0);  // float4 arrays of the staged records: A, B, F.., M
    constexpr int OFF_F = 2 * REC_STRIDE, OFF_M = (2 + FV) * REC_STRIDE;
    __shared__ float4 sRec[NREC][BATCH];
    __shared__ unsigned char sMask[BATCH];
    float4* const sA = sRec[0];
    float4* const sB = sRec[1];
    __shared__ unsigned char sList[TILE_PIX / 32][BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cam = blockIdx.z;
    const float tile_cx0 = (float)(blockIdx.x * TILE) + 0.5f, tile_cy0 = (float)(blockIdx.y * TILE) + 0.5f;
    const int tile_id = (cam * p.tile_h + blockIdx.y) * p.tile_w + blockIdx.x;
    int lx, ly;
    tile_pixel(tid, lx, ly);
    const int ix = blockIdx.x * TILE + lx, iy = blockIdx.y * TILE + ly;
    const float px = ix + 0.5f, py = iy + 0.5f;
    const bool inside = ix < p.width && iy < p.height;
    bool done = !inside;

    const int range_start = p.isect_offsets[tile_id];
    const int range_end = (tile_id == p.C * p.tile_h * p.tile_w - 1) ? (int)p.n_isects : p.isect_offsets[tile_id + 1];
    const int nb = (range_end - range_start + BATCH - 1) / BATCH;

    float T = 1.f;
    int cur_idx = 0;
    // the channels accumulate in pairs, one FFMA2 per pair (the weight is the instruction's broadcast operand): the same
    // IEEE fma per channel as the scalar form, half the issue slots -- the kernel is issue-bound (ncu r2: 80 % of the slots)
    constexpr int CP = (CH + 1) / 2;
    float2 acc2[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k) acc2[k] = make_float2(0.f, 0.f);

    const unsigned rec0 = smem_addr(&sRec[0][0]);
    const unsigned list0 = smem_addr(&sList[warp][0]);
    for (int b = 0; b < nb; ++b) {
        // barrier (previous batch fully consumed) + early exit when every pixel is finished
        if (__syncthreads_count(done) >= TILE_PIX) break;
        const int batch_start = range_start + b * BATCH;
        const int idx = batch_start + tid;
        if (idx < range_end) {
            const int g = p.flatten_ids[idx];
            const float2 m = p.means2d[g];
            const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
            const float opac = p.opacities[p.opac_shared ? g % p.N : g];
            const float a1 = 0.5f * LOG2E * ca, b1 = 0.5f * LOG2E * cb, c1 = 0.5f * LOG2E * cc;
            sA[tid] = make_float4(m.x, m.y, opac, a1);
            sB[tid] = make_float4(b1, c1, __int_as_float(g), 0.f);
            sMask[tid] = (unsigned char)patch_mask(m.x, m.y, opac, a1, 2.f * b1, c1, tile_cx0, tile_cy0);
            float f[FV * 4];
#pragma unroll
            for (int k = 0; k < FV * 4; ++k) f[k] = (k < CH) ? p.feat[(size_t)g * CH + k] : 0.f;
#pragma unroll
            for (int j = 0; j < FV; ++j) sRec[2 + j][tid] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            if (AFF) sRec[2 + FV][tid] = p.flow_affine[g];
        }
        __syncthreads();
        const int bs = min(BATCH, range_end - batch_start);
        // this warp's 8x4 patch only walks the Gaussians that can reach it
        const int n_list = build_warp_list(sMask, sList[warp], warp, lane, 0, bs);
        for (int li = 0; li < n_list && !done; ++li) {
            const int t = (int)lds_u8(list0 + li);
            const unsigned rec = rec0 + t * 16;
            const float4 a4 = lds128<0>(rec);
            const float2 b2 = lds64<REC_STRIDE>(rec);
            const GeomA ga = {a4.x, a4.y, a4.z, a4.w};
            const GeomB gb = {b2.x, b2.y, 0, 0.f};
            float dx, dy, u, v, vis, raw, alpha;
            if (!eval_alpha(ga, gb, px, py, dx, dy, u, v, vis, raw, alpha)) continue;
            const float next_T = T * (1.f - alpha);
            if (next_T <= T_STOP) {  // this Gaussian is not composited
                done = true;
                break;
            }
            const float w = alpha * T;
            float f[FV * 4];
            {
                const float4 q = lds128<OFF_F>(rec);
                f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
            }
            if (FV > 1) {
                const float4 q = lds128<OFF_F + REC_STRIDE>(rec);
                f[4 * (FV - 1)] = q.x; f[4 * (FV - 1) + 1] = q.y; f[4 * (FV - 1) + 2] = q.z; f[4 * (FV - 1) + 3] = q.w;
            }
            float e0 = 0.f, e1 = 0.f;
            if (AFF) {  // flow channels get + A (p - mu) = -A delta
                const float4 M = lds128<OFF_M>(rec);
                e0 = M.x * dx + M.y * dy;
                e1 = M.z * dx + M.w * dy;
            }
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                float f0 = f[2 * k], f1 = (2 * k + 1 < CH) ? f[2 * k + 1] : 0.f;
                if (AFF) {
                    f0 -= (2 * k == p.flow_ch0) ? e0 : (2 * k == p.flow_ch0 + 1) ? e1 : 0.f;
                    f1 -= (2 * k + 1 == p.flow_ch0) ? e0 : (2 * k + 1 == p.flow_ch0 + 1) ? e1 : 0.f;
                }
                acc2[k] = __ffma2_rn(make_float2(f0, f1), bc2(w), acc2[k]);
            }
            cur_idx = batch_start + t;
            T = next_T;
        }
    }
    if (inside) {
        const size_t pix = ((size_t)cam * p.height + iy) * p.width + ix;
        const float a_out = 1.f - T;
        p.alphas[pix] = a_out;
        const int n2 = CH - p.split;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            float v = (k & 1) ? acc2[k >> 1].y : acc2[k >> 1].x;
            if (p.backgrounds) v = fmaf(T, p.backgrounds[cam * CH + k], v);
            if (k == p.ed_channel) v = v / fmaxf(a_out, 1e-10f);
            if (k < p.split) p.render[pix * p.split + k] = v;
            else p.render2[pix * n2 + (k - p.split)] = v;
        }
        p.last_ids[pix] = cur_idx;
    }
}

// ---- two pixels per thread ------------------------------------------------------------------------------------------
// Same tile, same lists, same arithmetic as rasterize_fwd_kernel; a warp owns an 8x8 block and every lane the two pixels
// (x, y) and (x, y + 4) -- the two 8x4 patches of the culling mask, like the two-pixel backward kernel.  When both patches
// can be reached the pair is evaluated with packed FFMA2 / FMUL2 arithmetic (one instruction for both pixels, the
// Gaussian's record loaded once for 64 pixels); a patch the Gaussian cannot reach takes the scalar path of its half.
constexpr int FWD2_THREADS = 128;

template <int CH>
__device__ __forceinline__ void fwd_store_pixel(const RasterFwdParams& p, int cam, int ix, int iy, float T, const float (&acc)[CH],
                                                int cur_idx) {
    if (ix >= p.width || iy >= p.height) return;
    const size_t pix = ((size_t)cam * p.height + iy) * p.width + ix;
    const float a_out = 1.f - T;
    p.alphas[pix] = a_out;
    const int n2 = CH - p.split;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        float v = acc[k];
        if (p.backgrounds) v = fmaf(T, p.backgrounds[cam * CH + k], v);
        if (k == p.ed_channel) v = v / fmaxf(a_out, 1e-10f);
        if (k < p.split) p.render[pix * p.split + k] = v;
        else p.render2[pix * n2 + (k - p.split)] = v;
    }
    p.last_ids[pix] = cur_idx;
}

template <int CH>
__global__ void __launch_bounds__(FWD2_THREADS, 8) rasterize_fwd2_kernel(RasterFwdParams p) {
    pdl_wait();
    constexpr int FV = (CH + 3) / 4;
    constexpr int NREC = 2 + FV;
    constexpr int OFF_F = 2 * REC_STRIDE;
    constexpr int NW = FWD2_THREADS / 32;
    __shared__ float4 sRec[NREC][BATCH];
    __shared__ unsigned char sMask[BATCH];
    __shared__ unsigned short sList[NW][BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cam = blockIdx.z;
    const float tile_cx0 = (float)(blockIdx.x * TILE) + 0.5f, tile_cy0 = (float)(blockIdx.y * TILE) + 0.5f;
    const int tile_id = (cam * p.tile_h + blockIdx.y) * p.tile_w + blockIdx.x;
    const int bx = warp & 1, by = warp >> 1;
    const int ix = blockIdx.x * TILE + bx * 8 + (lane & 7);
    const int iy0 = blockIdx.y * TILE + by * 8 + (lane >> 3), iy1 = iy0 + 4;
    const float px = ix + 0.5f, py0 = iy0 + 0.5f;
    const float2 npy = make_float2(-py0, -(py0 + 4.f));
    const int w0 = by * 4 + bx;  // patch bits w0 (rows 0-3 of the block) and w0 + 2 (rows 4-7)
    bool done0 = !(ix < p.width && iy0 < p.height), done1 = !(ix < p.width && iy1 < p.height);

    const int range_start = p.isect_offsets[tile_id];
    const int range_end = (tile_id == p.C * p.tile_h * p.tile_w - 1) ? (int)p.n_isects : p.isect_offsets[tile_id + 1];
    const int nb = (range_end - range_start + BATCH - 1) / BATCH;

    float2 T = make_float2(1.f, 1.f);
    int cur0 = 0, cur1 = 0;
    float2 acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = make_float2(0.f, 0.f);

    const unsigned rec0 = smem_addr(&sRec[0][0]);
    const unsigned list0 = smem_addr(&sList[warp][0]);
    for (int b = 0; b < nb; ++b) {
        if (__syncthreads_count(done0 && done1) >= FWD2_THREADS) break;
        const int batch_start = range_start + b * BATCH;
#pragma unroll
        for (int h = 0; h < BATCH / FWD2_THREADS; ++h) {
            const int slot = tid + h * FWD2_THREADS;
            const int idx = batch_start + slot;
            if (idx < range_end) {
                const int g = p.flatten_ids[idx];
                const float2 m = p.means2d[g];
                const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
                const float opac = p.opacities[p.opac_shared ? g % p.N : g];
                const float a1 = 0.5f * LOG2E * ca, b1 = 0.5f * LOG2E * cb, c1 = 0.5f * LOG2E * cc;
                sRec[0][slot] = make_float4(m.x, m.y, opac, a1);
                sRec[1][slot] = make_float4(b1, c1, __int_as_float(g), 0.f);
                sMask[slot] = (unsigned char)patch_mask(m.x, m.y, opac, a1, 2.f * b1, c1, tile_cx0, tile_cy0);
                float f[FV * 4];
#pragma unroll
                for (int k = 0; k < FV * 4; ++k) f[k] = (k < CH) ? p.feat[(size_t)g * CH + k] : 0.f;
#pragma unroll
                for (int j = 0; j < FV; ++j) sRec[2 + j][slot] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
        }
        __syncthreads();
        const int bs = min(BATCH, range_end - batch_start);
        int n_list = 0;
        {
            const unsigned lt = (1u << lane) - 1;
#pragma unroll
            for (int c = 0; c < BATCH / 32; ++c) {
                const int t = c * 32 + lane;
                const unsigned mk = sMask[t];
                const unsigned hm = ((mk >> w0) & 1u) | (((mk >> (w0 + 2)) & 1u) << 1);
                const bool keep = t < bs && hm != 0;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) sList[warp][n_list + __popc(bal & lt)] = (unsigned short)(t | (hm << 8));
                n_list += __popc(bal);
            }
            __syncwarp();
        }
        for (int li = 0; li < n_list && !(done0 && done1); ++li) {
            unsigned e;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(list0 + 2 * li));
            const int t = e & 255;
            const unsigned rec = rec0 + t * 16;
            const float4 a4 = lds128<0>(rec);
            const float2 b2 = lds64<REC_STRIDE>(rec);
            const GeomA ga = {a4.x, a4.y, a4.z, a4.w};
            const GeomB gb = {b2.x, b2.y, 0, 0.f};
            const bool h0 = e & 0x100, h1 = e & 0x200;  // warp-uniform
            float2 w = make_float2(0.f, 0.f);
            if (h0 && h1) {
                float dx;
                float2 dy, u, v, vis, raw, alpha;
                bool ok0, ok1;
                eval_alpha_pair(ga, gb, px, npy, dx, dy, u, v, vis, raw, alpha, ok0, ok1);
                ok0 = ok0 && !done0;
                ok1 = ok1 && !done1;
                if (!(ok0 || ok1)) continue;
                const float2 next_T = __fmul2_rn(T, __ffma2_rn(alpha, bc2(-1.f), bc2(1.f)));
                if (ok0 && next_T.x <= T_STOP) { done0 = true; ok0 = false; }  // this Gaussian is not composited
                if (ok1 && next_T.y <= T_STOP) { done1 = true; ok1 = false; }
                const float2 wt = __fmul2_rn(alpha, T);
                w = make_float2(ok0 ? wt.x : 0.f, ok1 ? wt.y : 0.f);
                if (ok0) { T.x = next_T.x; cur0 = batch_start + t; }
                if (ok1) { T.y = next_T.y; cur1 = batch_start + t; }
            } else {
                float dx, dy, u, v, vis, raw, alpha;
                const bool ok = eval_alpha(ga, gb, px, h0 ? py0 : py0 + 4.f, dx, dy, u, v, vis, raw, alpha) && !(h0 ? done0 : done1);
                if (!ok) continue;
                const float Tj = h0 ? T.x : T.y;
                const float next_T = Tj * (1.f - alpha);
                if (next_T <= T_STOP) {
                    if (h0) done0 = true; else done1 = true;
                    continue;
                }
                if (h0) { w.x = alpha * Tj; T.x = next_T; cur0 = batch_start + t; }
                else { w.y = alpha * Tj; T.y = next_T; cur1 = batch_start + t; }
            }
            if (w.x == 0.f && w.y == 0.f) continue;
            float f[FV * 4];
            {
                const float4 q = lds128<OFF_F>(rec);
                f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
            }
            if (FV > 1) {
                const float4 q = lds128<OFF_F + REC_STRIDE>(rec);
                f[4 * (FV - 1)] = q.x; f[4 * (FV - 1) + 1] = q.y; f[4 * (FV - 1) + 2] = q.z; f[4 * (FV - 1) + 3] = q.w;
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) acc[k] = __ffma2_rn(bc2(f[k]), w, acc[k]);
        }
    }
    float a0[CH], a1v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) { a0[k] = acc[k].x; a1v[k] = acc[k].y; }
    fwd_store_pixel<CH>(p, cam, ix, iy0, T.x, a0, cur0);
    fwd_store_pixel<CH>(p, cam, ix, iy1, T.y, a1v, cur1);
}

template <int CH>
static int launch_raster_fwd(const RasterFwdParams& p, cudaStream_t st) {
    dim3 grid(p.tile_w, p.tile_h, p.C);
    if (p.flow_affine) {
        if (CH >= 2) {
            FG_LAUNCH((rasterize_fwd_kernel<(CH >= 2 ? CH : 2), true>), grid, TILE_PIX, 0, st, p);
        }
    } else if (g_fwd_two_pixels) {
        FG_LAUNCH((rasterize_fwd2_kernel<CH>), grid, FWD2_THREADS, 0, st, p);
    } else {
        FG_LAUNCH((rasterize_fwd_kernel<CH, false>), grid, TILE_PIX, 0, st, p);
    }
    return FG_OK;
}

}  // namespace fg

using namespace fg;

extern "C" int fg_rasterize_fwd(int C, int N, int CH, int width, int height, int tile_size, const float* means2d,
                                const float* conics, const float* feat, const float* opacities,
                                const float* backgrounds, const float* flow_affine, int flow_ch0, int split,
                                int ed_channel, int opac_shared, const int32_t* isect_offsets,
                                const int32_t* flatten_ids, int64_t n_isects, float* render, float* render2,
                                float* alphas, int32_t* last_ids, void* stream) {
    FG_REQUIRE(tile_size == TILE, "only tile_size=16 is supported (freegaussian_model.py:806)");
    FG_REQUIRE(C >= 1 && N >= 0 && width > 0 && height > 0, "bad C/N/width/height");
    FG_REQUIRE(CH >= 1 && CH <= FG_MAX_CHANNELS, "CH must be in 1..FG_MAX_CHANNELS");
    FG_REQUIRE(n_isects >= 0 && n_isects < (1ll << 31), "n_isects out of range");
    FG_REQUIRE(isect_offsets && render && alphas && last_ids, "NULL output/offset pointer");
    FG_REQUIRE(n_isects == 0 || (means2d && conics && feat && opacities && flatten_ids), "NULL input pointer");
    FG_REQUIRE(!flow_affine || (flow_ch0 >= 0 && flow_ch0 + 1 < CH), "flow_ch0 out of range");
    FG_REQUIRE(split >= 1 && split <= CH && (split == CH || render2), "bad split / render2");
    FG_REQUIRE(ed_channel >= -1 && ed_channel < CH, "ed_channel out of range");
    RasterFwdParams p;
    p.split = split; p.ed_channel = ed_channel; p.opac_shared = opac_shared; p.render2 = render2;
    p.C = C; p.N = N; p.width = width; p.height = height;
    p.tile_w = (width + TILE - 1) / TILE; p.tile_h = (height + TILE - 1) / TILE;
    p.means2d = (const float2*)means2d; p.conics = conics; p.feat = feat; p.opacities = opacities;
    p.backgrounds = backgrounds; p.flow_affine = (const float4*)flow_affine; p.flow_ch0 = flow_ch0;
    p.isect_offsets = isect_offsets; p.flatten_ids = flatten_ids; p.n_isects = n_isects;
    p.render = render; p.alphas = alphas; p.last_ids = last_ids;
    cudaStream_t st = (cudaStream_t)stream;
    switch (CH) {
        case 1: return launch_raster_fwd<1>(p, st);
        case 2: return launch_raster_fwd<2>(p, st);
        case 3: return launch_raster_fwd<3>(p, st);
        case 4: return launch_raster_fwd<4>(p, st);
        case 5: return launch_raster_fwd<5>(p, st);
        case 6: return launch_raster_fwd<6>(p, st);
        case 7: return launch_raster_fwd<7>(p, st);
        default: return launch_raster_fwd<8>(p, st);
    }
}
