// (3b) Per-tile alpha compositing, backward (pixel-parallel).  Two kernels: rasterize_bwd2_kernel (two pixels per
// thread, the default) and rasterize_bwd_kernel (one pixel per thread; carries the covariance-flow terms).
// Replaces gsplat rasterize_to_pixels_bwd behind the
// loss.backward() of the training step that calls freegaussian_model.py:847-868.
//
// Each thread replays its pixel back-to-front from last_ids with T = T_final, recovering
// T_i by division, and produces per-(pixel,Gaussian) partials for the CH feature channels,
// the conic (3), the 2-D mean (2), |d/dmean| (2, absgrad) and the opacity (1).  The only
// per-pixel state is (T_i, S_i) with S_i = sum_{j>i} w_j A_j, A_j = sum_k c_jk v_k: the
// gradient w.r.t. alpha_i is T_i A_i + (G - S_i)/(1 - alpha_i), so no per-channel buffer.
// The 8+CH partials are summed over the warp's 8x4 pixel patch with a TRANSPOSING butterfly
// (each stage halves the values a lane holds: 15 shuffles for 16 values instead of 80; ncu
// r1a showed the plain butterfly at ~2/3 of all instructions), after which 16 lanes hold one
// fully reduced value each and issue ONE atomic instruction per warp.  A warp whose 32 pixels
// all skip a Gaussian skips the whole reduction (the common case for small splats).
//
// Roofline: FP32 pipe + shuffle/atomic throughput; ~70 flop per evaluated pair at 6
// channels.
#include "rasterize_common.cuh"

namespace fg {

struct RasterBwdParams {
    int C, N, width, height, tile_w, tile_h;
    const float2* means2d;
    const float* conics;
    const float* feat;
    const float* opacities;
    const float* backgrounds;
    const float4* flow_affine;
    int flow_ch0;
    int split;       // v_render holds channels [0,split), v_render2 channels [split,CH)
    int ed_channel;  // "ED" channel (needs `render`), -1 = none
    int opac_shared; // opacities / v_opacities are [N] (index g % N)
    const int32_t* isect_offsets;
    const int32_t* flatten_ids;
    long long n_isects;
    const float* render;
    const float* alphas;
    const int32_t* last_ids;
    const float* v_render;
    const float* v_render2;
    const float* v_alphas;
    float* v_means2d;
    float* v_means2d_abs;
    float* v_conics;
    float* v_feat;
    float* v_opacities;
    float* v_flow_affine;
};

// Warp reduction of NV (16 or 32) partials per lane through a warp-private shared-memory
// transpose, 16 partials at a time: every lane stores them as a column (conflict-free STS), then
// lane l sums partial (l % 16) over lanes [16 (l / 16), +16) with four LDS.128 and the two halves
// are combined with one shuffle.  ~36 instructions for 16 values;
// the register butterfly it replaces took 85 (ncu r1g: 40 % of the kernel's instructions).
// Row stride 36 floats keeps the LDS.128 conflict-free.  On exit val[0] of lane l holds the
// warp-wide sum of partial l % 16 (both 16-lane halves hold every partial); for NV == 32, val[1]
// holds partial 16 + l % 16.  `wr` = shared address of buf[lane], `rd` = of buf[(l%16)*36 + (l/16)*16].
constexpr int RED_STRIDE = 36;
template <int NV, int NLIVE>
__device__ __forceinline__ void transpose_reduce(float (&val)[NV], unsigned wr, unsigned rd) {
    static_assert(NV == 16 || NV == 32, "NV must be 16 or 32");
    float out[NV / 16];
#pragma unroll
    for (int grp = 0; grp < NV / 16; ++grp) {
#define FG_STS_ROW(i) \
    if (grp * 16 + (i) < NLIVE) sts32<(i) * RED_STRIDE * 4>(wr, val[grp * 16 + (i)]);  // rows >= NLIVE: never consumed
        FG_STS_ROW(0) FG_STS_ROW(1) FG_STS_ROW(2) FG_STS_ROW(3) FG_STS_ROW(4) FG_STS_ROW(5) FG_STS_ROW(6) FG_STS_ROW(7)
        FG_STS_ROW(8) FG_STS_ROW(9) FG_STS_ROW(10) FG_STS_ROW(11) FG_STS_ROW(12) FG_STS_ROW(13) FG_STS_ROW(14) FG_STS_ROW(15)
#undef FG_STS_ROW
        __syncwarp();
        const float4 a = lds128<0>(rd), b = lds128<16>(rd), c = lds128<32>(rd), d = lds128<48>(rd);
        float s = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + ((c.x + c.y) + (c.z + c.w)) +
                  ((d.x + d.y) + (d.z + d.w));
        out[grp] = s + __shfl_xor_sync(0xffffffffu, s, 16);
        __syncwarp();  // buf is reused (next group / next Gaussian)
    }
#pragma unroll
    for (int grp = 0; grp < NV / 16; ++grp) val[grp] = out[grp];
}

template <int CH, bool AFF>
__global__ void __launch_bounds__(TILE_PIX, 4) rasterize_bwd_kernel(RasterBwdParams p) {
    pdl_wait();
    constexpr int FV = (CH + 3) / 4;
    constexpr int NVAL = CH + 8 + (AFF ? 4 : 0);  // partials per (pixel, Gaussian)
    constexpr int NV = NVAL <= 16 ? 16 : 32;
    constexpr int NREC = 2 + FV + (AFF ? 1 : 0);  // float4 arrays of the staged records: A, B, F.., M
    constexpr int OFF_F = 2 * REC_STRIDE, OFF_M = (2 + FV) * REC_STRIDE;
    __shared__ float4 sRec[NREC][BATCH];
    __shared__ unsigned char sMask[BATCH];
    __shared__ unsigned char sList[TILE_PIX / 32][BATCH];
    __shared__ __align__(16) float sRed[TILE_PIX / 32][16 * RED_STRIDE];
    float4* const sA = sRec[0];
    float4* const sB = sRec[1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float tile_cx0 = (float)(blockIdx.x * TILE) + 0.5f, tile_cy0 = (float)(blockIdx.y * TILE) + 0.5f;
    const int cam = blockIdx.z;
    const int tile_id = (cam * p.tile_h + blockIdx.y) * p.tile_w + blockIdx.x;
    int lx, ly;
    tile_pixel(tid, lx, ly);
    const int ix = blockIdx.x * TILE + lx, iy = blockIdx.y * TILE + ly;
    const float px = ix + 0.5f, py = iy + 0.5f;
    const bool inside = ix < p.width && iy < p.height;
    const size_t pix = ((size_t)cam * p.height + min(iy, p.height - 1)) * p.width + min(ix, p.width - 1);

    const int range_start = p.isect_offsets[tile_id];
    const int range_end = (tile_id == p.C * p.tile_h * p.tile_w - 1) ? (int)p.n_isects : p.isect_offsets[tile_id + 1];
    if (range_end <= range_start) return;

    // where this lane's reduced value goes: slot -> (array, elements per Gaussian, offset, scale)
    //   [0,CH) v_feat | CH..CH+2 v_conics | CH+3,4 v_means2d | CH+5,6 v_means2d_abs | CH+7 v_opacities | CH+8.. v_flow_affine
    // lanes 0..15 own partial `lane` (and lanes 16..31 partial 16 + (lane - 16) when NV == 32).
    // The constant factors of the conic partials (0.5, 1, 0.5) and of d sigma / d mean (2 ln 2, the
    // conic being staged pre-scaled) are applied here, once per reduced value instead of per pixel.
    float* out_base = nullptr;
    unsigned out_stride = 0;  // bytes per row; 0 = this lane owns no partial
    float out_scale = 1.f;
    bool out_is_opac = false;
    {
        const int slot = (NV == 32) ? lane : (lane < 16 ? lane : NV);  // NV: no partial
        if (slot < CH) { out_base = p.v_feat + slot; out_stride = CH * 4; }
        else if (slot < CH + 3) { out_base = p.v_conics + (slot - CH); out_stride = 12; out_scale = (slot == CH + 1) ? 1.f : 0.5f; }
        else if (slot < CH + 5) { out_base = p.v_means2d + (slot - CH - 3); out_stride = 8; out_scale = 2.f * LN2; }
        else if (slot < CH + 7) { if (p.v_means2d_abs) { out_base = p.v_means2d_abs + (slot - CH - 5); out_stride = 8; out_scale = 2.f * LN2; } }
        else if (slot < CH + 8) { out_base = p.v_opacities; out_stride = 4; out_is_opac = true; }
        else if (AFF && slot < CH + 12) { out_base = p.v_flow_affine + (slot - CH - 8); out_stride = 16; }
    }

    const float a_out = p.alphas[pix];
    const float T_final = 1.f - a_out;
    float T = T_final;
    float S = 0.f;  // sum over later list entries of w_j * A_j
    float v_out[CH];
    float v_alpha_out = (inside && p.v_alphas) ? p.v_alphas[pix] : 0.f;
    const int n2 = CH - p.split;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        float v = 0.f;
        if (inside) {
            if (k < p.split) v = p.v_render ? p.v_render[pix * p.split + k] : 0.f;
            else v = p.v_render2 ? p.v_render2[pix * n2 + (k - p.split)] : 0.f;
        }
        if (k == p.ed_channel) {
            // out = acc / max(alpha, 1e-10):  d/dacc = 1/alpha^,  d/dalpha = -out/alpha^ (0 under the clamp)
            const float inv = 1.f / fmaxf(a_out, 1e-10f);
            if (inside && a_out >= 1e-10f) v_alpha_out -= v * p.render[pix * p.split + k] * inv;
            v *= inv;
        }
        v_out[k] = v;
    }
    float bg_dot = 0.f;
    if (p.backgrounds) {
#pragma unroll
        for (int k = 0; k < CH; ++k) bg_dot += p.backgrounds[cam * CH + k] * v_out[k];
    }
    const float G = (v_alpha_out - bg_dot) * T_final;
    const int bin_final = inside ? p.last_ids[pix] : -1;
    const int warp_bin_final = __reduce_max_sync(0xffffffffu, bin_final);
    const int nb_all = (range_end - range_start + BATCH - 1) / BATCH;

    // explicit shared addresses of everything the inner loop touches
    const unsigned rec0 = smem_addr(&sRec[0][0]);
    const unsigned list0 = smem_addr(&sList[warp][0]);
    const unsigned red_wr = smem_addr(&sRed[warp][lane]);
    const unsigned red_rd = smem_addr(&sRed[warp][(lane & 15) * RED_STRIDE + (lane >> 4) * 16]);

    for (int b = 0; b < nb_all; ++b) {
        // batches run back to front; within a batch, smem slot t holds sorted index batch_end - t
        const int batch_end = range_end - 1 - BATCH * b;
        const int bs = min(BATCH, batch_end + 1 - range_start);
        // skip (uniformly) batches that lie entirely behind every pixel's last contributor
        const int need = __syncthreads_or(batch_end - bs + 1 <= warp_bin_final);
        if (!need) continue;
        const int idx = batch_end - tid;
        if (idx >= range_start) {
            const int g = p.flatten_ids[idx];
            const float2 m = p.means2d[g];
            const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
            const int go = p.opac_shared ? g % p.N : g;
            const float opac = p.opacities[go];
            const float a1 = 0.5f * LOG2E * ca, b1 = 0.5f * LOG2E * cb, c1 = 0.5f * LOG2E * cc;
            sA[tid] = make_float4(m.x, m.y, opac, a1);
            sMask[tid] = (unsigned char)patch_mask(m.x, m.y, opac, a1, 2.f * b1, c1, tile_cx0, tile_cy0);
            // .z: row of the per-(camera, Gaussian) gradients; .w: row of the opacity gradient (g, or
            // g % N when one opacity is shared by all cameras)
            sB[tid] = make_float4(b1, c1, __int_as_float(g), __int_as_float(go));
            float f[FV * 4];
#pragma unroll
            for (int k = 0; k < FV * 4; ++k) f[k] = (k < CH) ? p.feat[(size_t)g * CH + k] : 0.f;
#pragma unroll
            for (int j = 0; j < FV; ++j) sRec[2 + j][tid] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            if (AFF) sRec[2 + FV][tid] = p.flow_affine[g];
        }
        __syncthreads();
        // this warp's 8x4 patch only walks the Gaussians that can reach it (slot t <-> index batch_end - t)
        const int n_list = build_warp_list(sMask, sList[warp], warp, lane, max(0, batch_end - warp_bin_final), bs);
        const int t_lim = batch_end - bin_final;  // slots below it lie behind this pixel's last contributor
        for (int li = 0; li < n_list; ++li) {
            const int t = (int)lds_u8(list0 + li);
            const unsigned rec = rec0 + t * 16;
            const float4 a4 = lds128<0>(rec), b4 = lds128<REC_STRIDE>(rec);
            const GeomA ga = {a4.x, a4.y, a4.z, a4.w};
            const GeomB gb = {b4.x, b4.y, 0, 0.f};
            float dx, dy, u, v, vis, raw, alpha;
            bool valid = eval_alpha(ga, gb, px, py, dx, dy, u, v, vis, raw, alpha);
            valid = valid && (t >= t_lim);
            if (!__any_sync(0xffffffffu, valid)) continue;

            // Lanes whose pixel skips this Gaussian run the same arithmetic with alpha = vis = 0: every
            // partial then comes out as an exact zero and (T, S) are left unchanged (ra = 1), so no
            // per-partial zero-fill or branch is needed.
            if (!valid) { alpha = 0.f; vis = 0.f; }
            float val[NV];
#pragma unroll
            for (int k = CH + 8; k < NV; ++k) val[k] = 0.f;
            {
                float f[FV * 4];
                {
                    const float4 q = lds128<OFF_F>(rec);
                    f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
                }
                if (FV > 1) {
                    const float4 q = lds128<OFF_F + REC_STRIDE>(rec);
                    f[4 * (FV - 1)] = q.x; f[4 * (FV - 1) + 1] = q.y; f[4 * (FV - 1) + 2] = q.z; f[4 * (FV - 1) + 3] = q.w;
                }
                float4 M = make_float4(0.f, 0.f, 0.f, 0.f);
                float vo0 = 0.f, vo1 = 0.f;  // v_out of the two flow channels
                if (AFF) {
                    M = lds128<OFF_M>(rec);
                    const float e0 = M.x * dx + M.y * dy, e1 = M.z * dx + M.w * dy;
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        if (k == p.flow_ch0) { f[k] -= e0; vo0 = v_out[k]; }
                        if (k == p.flow_ch0 + 1) { f[k] -= e1; vo1 = v_out[k]; }
                    }
                }
                float ra;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.f - alpha));
                T *= ra;  // T_i, the transmittance in front of this Gaussian
                const float fac = alpha * T;
                float A = 0.f;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    A = fmaf(f[k], v_out[k], A);
                    val[k] = fac * v_out[k];
                }
                const float v_alpha = fmaf(T, A, (G - S) * ra);
                S = fmaf(fac, A, S);
                // gradient only flows through alpha where it is not clamped (opac * vis <= 0.999, where
                // alpha == opac * vis); on skipped lanes alpha = vis = 0 zero both products
                const bool gate = raw <= ALPHA_MAX;
                const float v_sigma = gate ? -alpha * v_alpha : 0.f;
                const float t1 = v_sigma * dx, t2 = v_sigma * dy;
                val[CH] = t1 * dx;      // x 0.5 in out_scale
                val[CH + 1] = t1 * dy;
                val[CH + 2] = t2 * dy;  // x 0.5
                const float hx = v_sigma * u, hy = v_sigma * v;  // x 2 ln2
                val[CH + 5] = fabsf(hx);
                val[CH + 6] = fabsf(hy);
                val[CH + 7] = gate ? vis * v_alpha : 0.f;
                if (AFF) {
                    // f_flow = feat - M delta: d/dM and d/ddelta of the composited flow
                    val[CH + 8] = -fac * vo0 * dx; val[CH + 9] = -fac * vo0 * dy;
                    val[CH + 10] = -fac * vo1 * dx; val[CH + 11] = -fac * vo1 * dy;
                    constexpr float inv_scale = 1.f / (2.f * LN2);
                    val[CH + 3] = fmaf(-fac * inv_scale, vo0 * M.x + vo1 * M.z, hx);
                    val[CH + 4] = fmaf(-fac * inv_scale, vo0 * M.y + vo1 * M.w, hy);
                } else {
                    val[CH + 3] = hx;
                    val[CH + 4] = hy;
                }
            }
            transpose_reduce<NV, NVAL>(val, red_wr, red_rd);
            if (out_stride) {
                const unsigned row = __float_as_uint(out_is_opac ? b4.w : b4.z);
                unsigned long long addr;  // out_base + row * stride
                asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(row), "r"(out_stride), "l"(out_base));
                const float sum = (NV == 32 && lane >= 16) ? val[1] : val[0] * out_scale;
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(sum) : "memory");
            }
        }
    }
}

// ---- two pixels per thread ----------------------------------------------------------------------
// The one-pixel kernel above is bound by shared-memory bandwidth: the transpose moves 30 of the ~36
// wavefronts a (warp, Gaussian) step costs (ncu r1y: 77 % of the LSU's wavefront rate, top stall
// short_scoreboard).  Here a warp owns an 8x8 block of the tile and every lane two pixels four rows
// apart (the two 8x4 patches of the culling mask); their partials are summed in registers -- the
// products fold into FFMAs -- before ONE transpose per 64 pixels.  A half whose patch bit is clear
// is skipped with a warp-uniform branch, so the per-patch culling loses nothing.
constexpr int BWD2_THREADS = 128;

template <int CH>
__device__ __forceinline__ void bwd_init_pixel(const RasterBwdParams& p, int cam, int ix, int iy, bool inside,
                                               float (&v_out)[CH], float& T, float& G, int& bin_final) {
    const size_t pix = ((size_t)cam * p.height + min(iy, p.height - 1)) * p.width + min(ix, p.width - 1);
    const float a_out = p.alphas[pix];
    const float T_final = 1.f - a_out;
    T = T_final;
    float v_alpha_out = (inside && p.v_alphas) ? p.v_alphas[pix] : 0.f;
    const int n2 = CH - p.split;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        float v = 0.f;
        if (inside) {
            if (k < p.split) v = p.v_render ? p.v_render[pix * p.split + k] : 0.f;
            else v = p.v_render2 ? p.v_render2[pix * n2 + (k - p.split)] : 0.f;
        }
        if (k == p.ed_channel) {
            const float inv = 1.f / fmaxf(a_out, 1e-10f);
            if (inside && a_out >= 1e-10f) v_alpha_out -= v * p.render[pix * p.split + k] * inv;
            v *= inv;
        }
        v_out[k] = v;
    }
    float bg_dot = 0.f;
    if (p.backgrounds) {
#pragma unroll
        for (int k = 0; k < CH; ++k) bg_dot += p.backgrounds[cam * CH + k] * v_out[k];
    }
    G = (v_alpha_out - bg_dot) * T_final;
    bin_final = inside ? p.last_ids[pix] : -1;
}

// partials of one pixel (half J of the thread's pixel pair) for one Gaussian, written (INIT) or added to val[]
template <int CH, bool INIT, int J, int NF>
__device__ __forceinline__ void bwd_pixel(bool valid, float alpha, float vis, float raw, float dx, float dy, float u,
                                          float v, const float (&f)[NF], const float2 (&V)[CH], float2& T2, float2& S2,
                                          float2 G2, float (&val)[16]) {
    float& T = J ? T2.y : T2.x;
    float& S = J ? S2.y : S2.x;
    const float G = J ? G2.y : G2.x;
    if (!valid) { alpha = 0.f; vis = 0.f; }
    float ra;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.f - alpha));
    T *= ra;
    const float fac = alpha * T;
    float A = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const float vo = J ? V[k].y : V[k].x;
        A = fmaf(f[k], vo, A);
        val[k] = INIT ? fac * vo : fmaf(fac, vo, val[k]);
    }
    const float v_alpha = fmaf(T, A, (G - S) * ra);
    S = fmaf(fac, A, S);
    const bool gate = raw <= ALPHA_MAX;
    const float v_sigma = gate ? -alpha * v_alpha : 0.f;
    const float t1 = v_sigma * dx, t2 = v_sigma * dy;
    const float hx = v_sigma * u, hy = v_sigma * v;
    const float vop = gate ? vis * v_alpha : 0.f;
    if (INIT) {
        val[CH] = t1 * dx; val[CH + 1] = t1 * dy; val[CH + 2] = t2 * dy;
        val[CH + 3] = hx; val[CH + 4] = hy; val[CH + 5] = fabsf(hx); val[CH + 6] = fabsf(hy); val[CH + 7] = vop;
    } else {
        val[CH] = fmaf(t1, dx, val[CH]); val[CH + 1] = fmaf(t1, dy, val[CH + 1]); val[CH + 2] = fmaf(t2, dy, val[CH + 2]);
        val[CH + 3] = fmaf(v_sigma, u, val[CH + 3]); val[CH + 4] = fmaf(v_sigma, v, val[CH + 4]);
        val[CH + 5] += fabsf(hx); val[CH + 6] += fabsf(hy); val[CH + 7] += vop;
    }
}

// ---- packed FP32: both pixels of the thread in one instruction (bc2 / eval_alpha_pair: rasterize_common.cuh) ----------
template <int CH, int NF>
__device__ __forceinline__ void bwd_pixel_pair(bool valid0, bool valid1, float2 alpha, float2 vis, float2 raw, float dx,
                                               float2 dy, float2 u, float2 v, const float (&f)[NF], const float2 (&V)[CH],
                                               float2& T, float2& S, float2 G, float (&val)[16]) {
    if (!valid0) { alpha.x = 0.f; vis.x = 0.f; }
    if (!valid1) { alpha.y = 0.f; vis.y = 0.f; }
    const float2 om = __ffma2_rn(alpha, bc2(-1.f), bc2(1.f));  // 1 - alpha
    float2 ra;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra.x) : "f"(om.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra.y) : "f"(om.y));
    T = __fmul2_rn(T, ra);
    const float2 fac = __fmul2_rn(alpha, T);
    float2 A = __fmul2_rn(bc2(f[0]), V[0]);
#pragma unroll
    for (int k = 1; k < CH; ++k) A = __ffma2_rn(bc2(f[k]), V[k], A);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const float2 pr = __fmul2_rn(fac, V[k]);
        val[k] = pr.x + pr.y;
    }
    const float2 gs = __ffma2_rn(S, bc2(-1.f), G);  // G - S
    const float2 v_alpha = __ffma2_rn(T, A, __fmul2_rn(gs, ra));
    S = __ffma2_rn(fac, A, S);
    const float2 nav = __fmul2_rn(alpha, v_alpha), vv = __fmul2_rn(vis, v_alpha);
    const bool g0 = raw.x <= ALPHA_MAX, g1 = raw.y <= ALPHA_MAX;
    const float2 v_sigma = make_float2(g0 ? -nav.x : 0.f, g1 ? -nav.y : 0.f);
    const float2 vop = make_float2(g0 ? vv.x : 0.f, g1 ? vv.y : 0.f);
    const float2 t2 = __fmul2_rn(v_sigma, dy);
    const float2 hx = __fmul2_rn(v_sigma, u), hy = __fmul2_rn(v_sigma, v);
    const float2 t2dy = __fmul2_rn(t2, dy);
    val[CH] = ((v_sigma.x + v_sigma.y) * dx) * dx;
    val[CH + 1] = (t2.x + t2.y) * dx;
    val[CH + 2] = t2dy.x + t2dy.y;
    val[CH + 3] = hx.x + hx.y;
    val[CH + 4] = hy.x + hy.y;
    val[CH + 5] = fabsf(hx.x) + fabsf(hx.y);
    val[CH + 6] = fabsf(hy.x) + fabsf(hy.y);
    val[CH + 7] = vop.x + vop.y;
}

template <int CH>
__global__ void __launch_bounds__(BWD2_THREADS, 6) rasterize_bwd2_kernel(RasterBwdParams p) {
    pdl_wait();
    constexpr int FV = (CH + 3) / 4;
    constexpr int NVAL = CH + 8;
    constexpr int NREC = 2 + FV;
    constexpr int OFF_F = 2 * REC_STRIDE;
    constexpr int NW = BWD2_THREADS / 32;
    static_assert(NVAL <= 16, "one transpose round");
    __shared__ float4 sRec[NREC][BATCH];
    __shared__ unsigned char sMask[BATCH];
    __shared__ unsigned short sList[NW][BATCH];
    __shared__ __align__(16) float sRed[NW][16 * RED_STRIDE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float tile_cx0 = (float)(blockIdx.x * TILE) + 0.5f, tile_cy0 = (float)(blockIdx.y * TILE) + 0.5f;
    const int cam = blockIdx.z;
    const int tile_id = (cam * p.tile_h + blockIdx.y) * p.tile_w + blockIdx.x;
    // warp -> 8x8 block (bx, by); lane -> column lane & 7, rows lane >> 3 and that + 4: the two 8x4
    // patches (mask bits w0, w0 + 2) of the forward kernel's tile_pixel() mapping
    const int bx = warp & 1, by = warp >> 1;
    const int ix = blockIdx.x * TILE + bx * 8 + (lane & 7);
    const int iy0 = blockIdx.y * TILE + by * 8 + (lane >> 3), iy1 = iy0 + 4;
    const float px = ix + 0.5f, py0 = iy0 + 0.5f;
    const int w0 = by * 4 + bx;

    const int range_start = p.isect_offsets[tile_id];
    const int range_end = (tile_id == p.C * p.tile_h * p.tile_w - 1) ? (int)p.n_isects : p.isect_offsets[tile_id + 1];
    if (range_end <= range_start) return;

    float* out_base = nullptr;
    unsigned out_stride = 0;
    float out_scale = 1.f;
    bool out_is_opac = false;
    {
        const int slot = lane < 16 ? lane : 16;
        if (slot < CH) { out_base = p.v_feat + slot; out_stride = CH * 4; }
        else if (slot < CH + 3) { out_base = p.v_conics + (slot - CH); out_stride = 12; out_scale = (slot == CH + 1) ? 1.f : 0.5f; }
        else if (slot < CH + 5) { out_base = p.v_means2d + (slot - CH - 3); out_stride = 8; out_scale = 2.f * LN2; }
        else if (slot < CH + 7) { if (p.v_means2d_abs) { out_base = p.v_means2d_abs + (slot - CH - 5); out_stride = 8; out_scale = 2.f * LN2; } }
        else if (slot < CH + 8) { out_base = p.v_opacities; out_stride = 4; out_is_opac = true; }
    }

    float2 V[CH];  // upstream gradients of the two pixels, channel by channel: (pixel 0, pixel 1)
    float2 T, G, S = make_float2(0.f, 0.f);
    int bin0, bin1;
    {
        float v_out0[CH], v_out1[CH];
        bwd_init_pixel<CH>(p, cam, ix, iy0, ix < p.width && iy0 < p.height, v_out0, T.x, G.x, bin0);
        bwd_init_pixel<CH>(p, cam, ix, iy1, ix < p.width && iy1 < p.height, v_out1, T.y, G.y, bin1);
#pragma unroll
        for (int k = 0; k < CH; ++k) V[k] = make_float2(v_out0[k], v_out1[k]);
    }
    const float2 npy = make_float2(-py0, -(py0 + 4.f));
    const int warp_bin_final = __reduce_max_sync(0xffffffffu, max(bin0, bin1));
    const int nb_all = (range_end - range_start + BATCH - 1) / BATCH;

    const unsigned rec0 = smem_addr(&sRec[0][0]);
    const unsigned list0 = smem_addr(&sList[warp][0]);
    const unsigned red_wr = smem_addr(&sRed[warp][lane]);
    const unsigned red_rd = smem_addr(&sRed[warp][(lane & 15) * RED_STRIDE + (lane >> 4) * 16]);

    for (int b = 0; b < nb_all; ++b) {
        const int batch_end = range_end - 1 - BATCH * b;
        const int bs = min(BATCH, batch_end + 1 - range_start);
        const int need = __syncthreads_or(batch_end - bs + 1 <= warp_bin_final);
        if (!need) continue;
#pragma unroll
        for (int h = 0; h < BATCH / BWD2_THREADS; ++h) {
            const int slot = tid + h * BWD2_THREADS;
            const int idx = batch_end - slot;
            if (idx >= range_start) {
                const int g = p.flatten_ids[idx];
                const float2 m = p.means2d[g];
                const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
                const int go = p.opac_shared ? g % p.N : g;
                const float opac = p.opacities[go];
                const float a1 = 0.5f * LOG2E * ca, b1 = 0.5f * LOG2E * cb, c1 = 0.5f * LOG2E * cc;
                sRec[0][slot] = make_float4(m.x, m.y, opac, a1);
                sMask[slot] = (unsigned char)patch_mask(m.x, m.y, opac, a1, 2.f * b1, c1, tile_cx0, tile_cy0);
                sRec[1][slot] = make_float4(b1, c1, __int_as_float(g), __int_as_float(go));
                float f[FV * 4];
#pragma unroll
                for (int k = 0; k < FV * 4; ++k) f[k] = (k < CH) ? p.feat[(size_t)g * CH + k] : 0.f;
#pragma unroll
                for (int j = 0; j < FV; ++j) sRec[2 + j][slot] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
        }
        __syncthreads();
        // this warp's list: slot | (which halves can be reached) << 8, in order
        int n_list = 0;
        {
            const int t_min = max(0, batch_end - warp_bin_final);
            const unsigned lt = (1u << lane) - 1;
#pragma unroll
            for (int c = 0; c < BATCH / 32; ++c) {
                const int t = c * 32 + lane;
                const unsigned mk = sMask[t];
                const unsigned hm = ((mk >> w0) & 1u) | (((mk >> (w0 + 2)) & 1u) << 1);
                const bool keep = t >= t_min && t < bs && hm != 0;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) sList[warp][n_list + __popc(bal & lt)] = (unsigned short)(t | (hm << 8));
                n_list += __popc(bal);
            }
            __syncwarp();
        }
        const int t_lim0 = batch_end - bin0, t_lim1 = batch_end - bin1;
        for (int li = 0; li < n_list; ++li) {
            unsigned e;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(list0 + 2 * li));
            const int t = e & 255;
            const unsigned rec = rec0 + t * 16;
            const float4 a4 = lds128<0>(rec), b4 = lds128<REC_STRIDE>(rec);
            const GeomA ga = {a4.x, a4.y, a4.z, a4.w};
            const GeomB gb = {b4.x, b4.y, 0, 0.f};
            const bool h0 = e & 0x100, h1 = e & 0x200;  // warp-uniform
            float val[16];
            if (h0 && h1) {
                // both halves reachable: packed arithmetic over the pixel pair
                float dx;
                float2 dy, u, v, vis, raw, alpha;
                bool valid0, valid1;
                eval_alpha_pair(ga, gb, px, npy, dx, dy, u, v, vis, raw, alpha, valid0, valid1);
                valid0 = valid0 && (t >= t_lim0);
                valid1 = valid1 && (t >= t_lim1);
                if (!__any_sync(0xffffffffu, valid0 || valid1)) continue;
                float f[FV * 4];
                {
                    const float4 q = lds128<OFF_F>(rec);
                    f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
                }
                if (FV > 1) {
                    const float4 q = lds128<OFF_F + REC_STRIDE>(rec);
                    f[4 * (FV - 1)] = q.x; f[4 * (FV - 1) + 1] = q.y; f[4 * (FV - 1) + 2] = q.z; f[4 * (FV - 1) + 3] = q.w;
                }
                bwd_pixel_pair<CH>(valid0, valid1, alpha, vis, raw, dx, dy, u, v, f, V, T, S, G, val);
            } else {
                float dx, dy = 0.f, u = 0.f, v = 0.f, vis = 0.f, raw = 0.f, alpha = 0.f;
                bool valid;
                if (h0) valid = eval_alpha(ga, gb, px, py0, dx, dy, u, v, vis, raw, alpha) && (t >= t_lim0);
                else valid = eval_alpha(ga, gb, px, py0 + 4.f, dx, dy, u, v, vis, raw, alpha) && (t >= t_lim1);
                if (!__any_sync(0xffffffffu, valid)) continue;
                float f[FV * 4];
                {
                    const float4 q = lds128<OFF_F>(rec);
                    f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
                }
                if (FV > 1) {
                    const float4 q = lds128<OFF_F + REC_STRIDE>(rec);
                    f[4 * (FV - 1)] = q.x; f[4 * (FV - 1) + 1] = q.y; f[4 * (FV - 1) + 2] = q.z; f[4 * (FV - 1) + 3] = q.w;
                }
                if (h0) bwd_pixel<CH, true, 0>(valid, alpha, vis, raw, dx, dy, u, v, f, V, T, S, G, val);
                else bwd_pixel<CH, true, 1>(valid, alpha, vis, raw, dx, dy, u, v, f, V, T, S, G, val);
            }
#pragma unroll
            for (int k = NVAL; k < 16; ++k) val[k] = 0.f;
            transpose_reduce<16, NVAL>(val, red_wr, red_rd);
            if (out_stride) {
                const unsigned row = __float_as_uint(out_is_opac ? b4.w : b4.z);
                unsigned long long addr;
                asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(row), "r"(out_stride), "l"(out_base));
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(val[0] * out_scale) : "memory");
            }
        }
    }
}

template <int CH>
static int launch_raster_bwd(const RasterBwdParams& p, cudaStream_t st) {
    dim3 grid(p.tile_w, p.tile_h, p.C);
    if (p.flow_affine) {
        if (CH >= 2) {
            FG_LAUNCH((rasterize_bwd_kernel<(CH >= 2 ? CH : 2), true>), grid, TILE_PIX, 0, st, p);
        }
    } else {
        FG_LAUNCH((rasterize_bwd2_kernel<CH>), grid, BWD2_THREADS, 0, st, p);
    }
    return FG_OK;
}

}  // namespace fg

using namespace fg;

extern "C" int fg_rasterize_bwd(int C, int N, int CH, int width, int height, int tile_size, const float* means2d,
                                const float* conics, const float* feat, const float* opacities,
                                const float* backgrounds, const float* flow_affine, int flow_ch0, int split,
                                int ed_channel, int opac_shared, const int32_t* isect_offsets,
                                const int32_t* flatten_ids, int64_t n_isects, const float* render,
                                const float* alphas, const int32_t* last_ids, const float* v_render,
                                const float* v_render2, const float* v_alphas, float* v_means2d,
                                float* v_means2d_abs, float* v_conics, float* v_feat, float* v_opacities,
                                float* v_flow_affine, void* stream) {
    FG_REQUIRE(tile_size == TILE, "only tile_size=16 is supported (freegaussian_model.py:806)");
    FG_REQUIRE(C >= 1 && N >= 0 && width > 0 && height > 0, "bad C/N/width/height");
    FG_REQUIRE(CH >= 1 && CH <= FG_MAX_CHANNELS, "CH must be in 1..FG_MAX_CHANNELS");
    FG_REQUIRE(n_isects >= 0 && n_isects < (1ll << 31), "n_isects out of range");
    if (n_isects == 0) return FG_OK;
    FG_REQUIRE(means2d && conics && feat && opacities && isect_offsets && flatten_ids && alphas && last_ids,
               "NULL input pointer");
    FG_REQUIRE(split >= 1 && split <= CH, "bad split");
    FG_REQUIRE(ed_channel >= -1 && ed_channel < split && (ed_channel < 0 || render), "ed_channel needs render and must be < split");
    FG_REQUIRE(v_means2d && v_conics && v_feat && v_opacities, "NULL gradient output pointer");
    FG_REQUIRE(!flow_affine || (flow_ch0 >= 0 && flow_ch0 + 1 < CH && v_flow_affine), "bad flow_affine arguments");
    RasterBwdParams p;
    p.C = C; p.N = N; p.width = width; p.height = height;
    p.tile_w = (width + TILE - 1) / TILE; p.tile_h = (height + TILE - 1) / TILE;
    p.means2d = (const float2*)means2d; p.conics = conics; p.feat = feat; p.opacities = opacities;
    p.backgrounds = backgrounds; p.flow_affine = (const float4*)flow_affine; p.flow_ch0 = flow_ch0;
    p.isect_offsets = isect_offsets; p.flatten_ids = flatten_ids; p.n_isects = n_isects;
    p.alphas = alphas; p.last_ids = last_ids; p.v_render = v_render; p.v_alphas = v_alphas;
    p.split = split; p.ed_channel = ed_channel; p.opac_shared = opac_shared; p.render = render; p.v_render2 = v_render2;
    p.v_means2d = v_means2d; p.v_means2d_abs = v_means2d_abs; p.v_conics = v_conics; p.v_feat = v_feat;
    p.v_opacities = v_opacities; p.v_flow_affine = v_flow_affine;
    cudaStream_t st = (cudaStream_t)stream;
    switch (CH) {
        case 1: return launch_raster_bwd<1>(p, st);
        case 2: return launch_raster_bwd<2>(p, st);
        case 3: return launch_raster_bwd<3>(p, st);
        case 4: return launch_raster_bwd<4>(p, st);
        case 5: return launch_raster_bwd<5>(p, st);
        case 6: return launch_raster_bwd<6>(p, st);
        case 7: return launch_raster_bwd<7>(p, st);
        default: return launch_raster_bwd<8>(p, st);
    }
}
