// (3b') Per-tile alpha compositing, backward -- Gaussian-parallel variant.
// Same result as rasterize_bwd.cu with a different mapping that
// removes the per-(pixel,Gaussian) warp reductions, the dominant cost of the pixel-parallel
// kernel (ncu r1a: 70 SHFL + 14 atomics per warp per Gaussian, FMA pipe 25 % busy).
//
// For pixel p and list entry i with alpha_i, T_i = prod_{j<i}(1-alpha_j), w_i = alpha_i T_i,
// A_i = sum_k c_ik v_k (v = dL/dout of the pixel) and P_i = sum_{j<i} w_j A_j:
//     dL/dalpha_i = T_i A_i + (G - (Q - P_i - w_i A_i)) / (1 - alpha_i),
//     Q = sum_k v_k (out_k - T_final bg_k),   G = (dL/dalpha_out - sum_k bg_k v_k) T_final,
// so the only per-pixel state that has to flow along the list is the pair (T_i, P_i).
//
// Per batch of 256 list entries the CTA (8 warps, 256 threads) does
//   phase 1 (pixel-parallel replay): thread = pixel walks the batch front to back updating
//           (T, P) and stores a checkpoint at each 32-entry bucket start;
//   phase 2 (Gaussian-parallel): warp = bucket, lane = Gaussian.  The 256 pixels stream through
//           the warp systolically -- at step s lane l handles pixel s-l and hands (T, P) to lane
//           l+1 with one shuffle pair -- while every lane accumulates ITS Gaussian's 8+CH gradient
//           values in registers over all pixels.  One atomic per value per (tile, Gaussian).
//
// Roofline: FP32 pipe; ~24 flop (replay) + ~46 flop (main) per evaluated pair at 6 channels.
#include "rasterize_common.cuh"

namespace fg {

struct RasterBwdGpParams {
    int C, N, width, height, tile_w, tile_h;
    const float2* means2d;
    const float* conics;
    const float* feat;
    const float* opacities;
    const float* backgrounds;
    const int32_t* isect_offsets;
    const int32_t* flatten_ids;
    long long n_isects;
    const float* render;
    const float* alphas;
    const int32_t* last_ids;
    const float* v_render;
    const float* v_alphas;
    float* v_means2d;
    float* v_means2d_abs;
    float* v_conics;
    float* v_feat;
    float* v_opacities;
};

constexpr int BUCKET = 32;
constexpr int NBUCKET = BATCH / BUCKET;  // 8 = warps per CTA

template <int CH>
__global__ void __launch_bounds__(TILE_PIX) rasterize_bwd_gp_kernel(RasterBwdGpParams p) {
    constexpr int FV = (CH + 3) / 4;      // float4 per Gaussian: features
    constexpr int PV = (CH + 2 + 3) / 4;  // float4 per pixel: v_out[CH], G, Q
    __shared__ float4 sA[BATCH];
    __shared__ float4 sB[BATCH];
    __shared__ float4 sF[FV][BATCH];
    __shared__ float2 sCk[NBUCKET][TILE_PIX];
    __shared__ float4 sPix[PV][TILE_PIX];
    __shared__ int sLast[TILE_PIX];
    __shared__ unsigned char sList[NBUCKET][TILE_PIX];  // per bucket: ids of the pixels that reach it
    __shared__ int sMaxLast;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cam = blockIdx.z;
    const int tile_id = (cam * p.tile_h + blockIdx.y) * p.tile_w + blockIdx.x;
    const int range_start = p.isect_offsets[tile_id];
    const int range_end = (tile_id == p.C * p.tile_h * p.tile_w - 1) ? (int)p.n_isects : p.isect_offsets[tile_id + 1];
    if (range_end <= range_start) return;

    // ---- per-pixel data (thread = pixel tid)
    // pixel id = row-major index inside the tile (cheap to turn back into coordinates in phase 2)
    const int lx = tid & (TILE - 1), ly = tid >> 4;
    const int ix = blockIdx.x * TILE + lx, iy = blockIdx.y * TILE + ly;
    const float px = ix + 0.5f, py = iy + 0.5f;
    const bool inside = ix < p.width && iy < p.height;
    const size_t pix = ((size_t)cam * p.height + min(iy, p.height - 1)) * p.width + min(ix, p.width - 1);
    float v_out[CH];
    float Q = 0.f, bg_dot = 0.f;
    const float T_final = inside ? 1.f - p.alphas[pix] : 1.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        v_out[k] = inside ? p.v_render[pix * CH + k] : 0.f;
        float o = inside ? p.render[pix * CH + k] : 0.f;
        if (p.backgrounds) {
            const float bg = p.backgrounds[cam * CH + k];
            o -= T_final * bg;
            bg_dot += bg * v_out[k];
        }
        Q += v_out[k] * o;
    }
    const float G = (((inside && p.v_alphas) ? p.v_alphas[pix] : 0.f) - bg_dot) * T_final;
    const int my_last = inside ? p.last_ids[pix] : -1;
    {
        float pd[PV * 4];
#pragma unroll
        for (int k = 0; k < PV * 4; ++k) pd[k] = 0.f;
#pragma unroll
        for (int k = 0; k < CH; ++k) pd[k] = v_out[k];
        pd[CH] = G;
        pd[CH + 1] = Q;
#pragma unroll
        for (int j = 0; j < PV; ++j) sPix[j][tid] = make_float4(pd[4 * j], pd[4 * j + 1], pd[4 * j + 2], pd[4 * j + 3]);
        sLast[tid] = my_last;
    }
    if (tid == 0) sMaxLast = -1;
    __syncthreads();
    const int warp_last = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0) atomicMax(&sMaxLast, warp_last);
    __syncthreads();
    const int tile_last = min(sMaxLast, range_end - 1);
    if (tile_last < range_start) return;
    const int nb = (tile_last - range_start) / BATCH + 1;

    float T = 1.f, P = 0.f;  // replay state of this thread's pixel
    for (int b = 0; b < nb; ++b) {
        const int batch_start = range_start + b * BATCH;
        const int bs = min(BATCH, range_end - batch_start);
        __syncthreads();  // previous batch fully consumed
        if (tid < bs) {
            const int g = p.flatten_ids[batch_start + tid];
            const float2 m = p.means2d[g];
            const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
            sA[tid] = make_float4(m.x, m.y, p.opacities[g], 0.5f * LOG2E * ca);
            sB[tid] = make_float4(LOG2E * cb, 0.5f * LOG2E * cc, __int_as_float(g), 0.f);
            float f[FV * 4];
#pragma unroll
            for (int k = 0; k < FV * 4; ++k) f[k] = (k < CH) ? p.feat[(size_t)g * CH + k] : 0.f;
#pragma unroll
            for (int j = 0; j < FV; ++j) sF[j][tid] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        __syncthreads();

        // ---- phase 1: replay (thread = pixel), checkpoint (T,P) at every bucket start
        {
            const int t_stop = min(bs, warp_last - batch_start + 1);  // nothing beyond the warp's last contributor
#pragma unroll 1
            for (int k = 0; k < NBUCKET; ++k) {
                sCk[k][tid] = make_float2(T, P);
                const int t1 = min(t_stop, (k + 1) * BUCKET);
                for (int t = k * BUCKET; t < t1; ++t) {
                    if (batch_start + t > my_last) continue;
                    const float4 a4 = sA[t], b4 = sB[t];
                    const GeomA ga = {a4.x, a4.y, a4.z, a4.w};
                    const GeomB gb = {b4.x, b4.y, 0, 0.f};
                    float dx, dy, vis, alpha;
                    if (!eval_alpha(ga, gb, px, py, dx, dy, vis, alpha)) continue;
                    float f[FV * 4];
#pragma unroll
                    for (int j = 0; j < FV; ++j) {
                        const float4 v = sF[j][t];
                        f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
                    }
                    float A = 0.f;
#pragma unroll
                    for (int k = 0; k < CH; ++k) A = fmaf(f[k], v_out[k], A);
                    P = fmaf(alpha * T, A, P);
                    T *= (1.f - alpha);
                }
            }
        }
        __syncthreads();

        // ---- phase 2: warp = bucket, lane = Gaussian; pixels stream through the warp
        const int slot = warp * BUCKET + lane;
        const int bucket_first = batch_start + warp * BUCKET;
        if (warp * BUCKET < bs && bucket_first <= tile_last) {
            const bool has_g = slot < bs;
            const int gidx = batch_start + slot;
            const float4 a4 = has_g ? sA[slot] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b4 = has_g ? sB[slot] : make_float4(0.f, 0.f, 0.f, 0.f);
            const GeomA ga = {a4.x, a4.y, a4.z, a4.w};
            const GeomB gb = {b4.x, b4.y, 0, 0.f};
            float c[FV * 4];
#pragma unroll
            for (int j = 0; j < FV; ++j) {
                const float4 v = has_g ? sF[j][slot] : make_float4(0.f, 0.f, 0.f, 0.f);
                c[4 * j] = v.x; c[4 * j + 1] = v.y; c[4 * j + 2] = v.z; c[4 * j + 3] = v.w;
            }
            float acc_c[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) acc_c[k] = 0.f;
            float acc_ca = 0.f, acc_cb = 0.f, acc_cc = 0.f, acc_x = 0.f, acc_y = 0.f, acc_ax = 0.f, acc_ay = 0.f,
                  acc_op = 0.f;
            bool touched = false;
            float T_out = 1.f, P_out = 0.f;
            // pixels whose last contributor lies at or beyond this bucket, in pixel order
            int n_act = 0;
            {
                const unsigned lt = (1u << lane) - 1;
#pragma unroll
                for (int i = 0; i < TILE_PIX / 32; ++i) {
                    const int q = i * 32 + lane;
                    const bool act = sLast[q] >= bucket_first;
                    const unsigned m = __ballot_sync(0xffffffffu, act);
                    if (act) sList[warp][n_act + __popc(m & lt)] = (unsigned char)q;
                    n_act += __popc(m);
                }
                __syncwarp();
            }
            const float tile_x0 = (float)(blockIdx.x * TILE) + 0.5f, tile_y0 = (float)(blockIdx.y * TILE) + 0.5f;
            const int n_steps = n_act + BUCKET - 1;
#pragma unroll 2
            for (int s = 0; s < n_steps; ++s) {
                float T_in = __shfl_up_sync(0xffffffffu, T_out, 1);
                float P_in = __shfl_up_sync(0xffffffffu, P_out, 1);
                const int si = s - lane;  // position of this lane's pixel in the active list
                const bool live = si >= 0 && si < n_act;
                const int q = live ? (int)sList[warp][si] : 0;
                if (lane == 0 && live) {
                    const float2 ck = sCk[warp][q];
                    T_in = ck.x;
                    P_in = ck.y;
                }
                T_out = T_in;
                P_out = P_in;
                if (!live || !has_g) continue;
                if (gidx > sLast[q]) continue;
                const float fx = tile_x0 + (float)(q & (TILE - 1)), fy = tile_y0 + (float)(q >> 4);
                float dx, dy, vis, alpha;
                if (!eval_alpha(ga, gb, fx, fy, dx, dy, vis, alpha)) continue;
                float pd[PV * 4];
#pragma unroll
                for (int j = 0; j < PV; ++j) {
                    const float4 v = sPix[j][q];
                    pd[4 * j] = v.x; pd[4 * j + 1] = v.y; pd[4 * j + 2] = v.z; pd[4 * j + 3] = v.w;
                }
                float A = 0.f;
#pragma unroll
                for (int k = 0; k < CH; ++k) A = fmaf(c[k], pd[k], A);
                const float w = alpha * T_in;
                const float wA = w * A;
                float ra;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.f - alpha));
                const float v_alpha = fmaf(T_in, A, (pd[CH] - (pd[CH + 1] - P_in - wA)) * ra);
#pragma unroll
                for (int k = 0; k < CH; ++k) acc_c[k] = fmaf(w, pd[k], acc_c[k]);
                if (a4.z * vis <= ALPHA_MAX) {
                    const float v_sigma = -a4.z * vis * v_alpha;
                    acc_ca = fmaf(0.5f * v_sigma * dx, dx, acc_ca);
                    acc_cb = fmaf(v_sigma * dx, dy, acc_cb);
                    acc_cc = fmaf(0.5f * v_sigma * dy, dy, acc_cc);
                    const float vs = v_sigma * LN2;
                    const float gx = vs * (2.f * a4.w * dx + b4.x * dy);
                    const float gy = vs * (b4.x * dx + 2.f * b4.y * dy);
                    acc_x += gx; acc_y += gy;
                    acc_ax += fabsf(gx); acc_ay += fabsf(gy);
                    acc_op = fmaf(vis, v_alpha, acc_op);
                }
                touched = true;
                T_out = T_in * (1.f - alpha);
                P_out = P_in + wA;
            }
            if (touched) {
                const size_t g = (size_t)__float_as_int(b4.z);
#pragma unroll
                for (int k = 0; k < CH; ++k) atomicAdd(p.v_feat + g * CH + k, acc_c[k]);
                atomicAdd(p.v_conics + 3 * g, acc_ca);
                atomicAdd(p.v_conics + 3 * g + 1, acc_cb);
                atomicAdd(p.v_conics + 3 * g + 2, acc_cc);
                atomicAdd(p.v_means2d + 2 * g, acc_x);
                atomicAdd(p.v_means2d + 2 * g + 1, acc_y);
                if (p.v_means2d_abs) {
                    atomicAdd(p.v_means2d_abs + 2 * g, acc_ax);
                    atomicAdd(p.v_means2d_abs + 2 * g + 1, acc_ay);
                }
                atomicAdd(p.v_opacities + g, acc_op);
            }
        }
    }
}

template <int CH>
static int launch_raster_bwd_gp(const RasterBwdGpParams& p, cudaStream_t st) {
    dim3 grid(p.tile_w, p.tile_h, p.C);
    FG_LAUNCH((rasterize_bwd_gp_kernel<CH>), grid, TILE_PIX, 0, st, p);
    return FG_OK;
}

}  // namespace fg

using namespace fg;

extern "C" int fg_rasterize_bwd_gp(int C, int N, int CH, int width, int height, int tile_size, const float* means2d,
                                   const float* conics, const float* feat, const float* opacities,
                                   const float* backgrounds, const int32_t* isect_offsets,
                                   const int32_t* flatten_ids, int64_t n_isects, const float* render,
                                   const float* alphas, const int32_t* last_ids, const float* v_render,
                                   const float* v_alphas, float* v_means2d, float* v_means2d_abs, float* v_conics,
                                   float* v_feat, float* v_opacities, void* stream) {
    FG_REQUIRE(tile_size == TILE, "only tile_size=16 is supported (freegaussian_model.py:806)");
    FG_REQUIRE(C >= 1 && N >= 0 && width > 0 && height > 0, "bad C/N/width/height");
    FG_REQUIRE(CH >= 1 && CH <= FG_MAX_CHANNELS, "CH must be in 1..FG_MAX_CHANNELS");
    FG_REQUIRE(n_isects >= 0 && n_isects < (1ll << 31), "n_isects out of range");
    if (n_isects == 0) return FG_OK;
    FG_REQUIRE(means2d && conics && feat && opacities && isect_offsets && flatten_ids && render && alphas &&
                   last_ids && v_render,
               "NULL input pointer");
    FG_REQUIRE(v_means2d && v_conics && v_feat && v_opacities, "NULL gradient output pointer");
    RasterBwdGpParams p;
    p.C = C; p.N = N; p.width = width; p.height = height;
    p.tile_w = (width + TILE - 1) / TILE; p.tile_h = (height + TILE - 1) / TILE;
    p.means2d = (const float2*)means2d; p.conics = conics; p.feat = feat; p.opacities = opacities;
    p.backgrounds = backgrounds; p.isect_offsets = isect_offsets; p.flatten_ids = flatten_ids; p.n_isects = n_isects;
    p.render = render; p.alphas = alphas; p.last_ids = last_ids; p.v_render = v_render; p.v_alphas = v_alphas;
    p.v_means2d = v_means2d; p.v_means2d_abs = v_means2d_abs; p.v_conics = v_conics; p.v_feat = v_feat;
    p.v_opacities = v_opacities;
    cudaStream_t st = (cudaStream_t)stream;
    switch (CH) {
        case 1: return launch_raster_bwd_gp<1>(p, st);
        case 2: return launch_raster_bwd_gp<2>(p, st);
        case 3: return launch_raster_bwd_gp<3>(p, st);
        case 4: return launch_raster_bwd_gp<4>(p, st);
        case 5: return launch_raster_bwd_gp<5>(p, st);
        case 6: return launch_raster_bwd_gp<6>(p, st);
        case 7: return launch_raster_bwd_gp<7>(p, st);
        default: return launch_raster_bwd_gp<8>(p, st);
    }
}
