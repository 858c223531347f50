// Fused blend + clamp + L1 + SSIM loss, forward and backward: the
// image-space step that immediately follows the render call.
//   pred = clamp(render[..., :3] + (1 - alpha) * bg, 0, 1)                     freegaussian_model.py:876-877
//   loss = (1 - l) * mean|gt - pred| + l * (1 - SSIM(gt, pred))                 freegaussian_model.py:965-981
//   SSIM = pytorch_msssim (11-tap Gaussian, sigma 1.5, separable valid convolution, K = (0.01, 0.03))
//
// Forward: one CTA per 32x16 pixel block and channel.  The (32+10)x(16+10) halo of pred and gt is
// built in shared memory (pred on the fly from render/alpha/bg), filtered horizontally for the five
// moments (x, y, xx, yy, xy), then vertically per output pixel.  The kernel writes the three partial
// maps dS/d(mu_x), dS/d(E xx), dS/d(E xy) -- pre-scaled with -l / (3 * valid pixels) -- and reduces
// the L1 and SSIM sums.  Backward: the transposed (full) separable convolution of those maps,
// combined with the L1 sign and the clamp mask, chained to render and alpha.
//
// Roofline: HBM.  Forward reads render, alpha, gt once (+halo re-reads through L2) and writes 3 floats per
// channel per valid pixel; backward reads them back and writes v_render, v_alpha.  ~50 floats of
// traffic per pixel against ~15 full-image elementwise/conv passes in the torch formulation.
#include <algorithm>

#include "common.cuh"

namespace fg {

constexpr int LW = 11, LR = 5;           // window size / radius
constexpr int LTX = 32, LTY = 16;        // output block
constexpr int LHX = LTX + LW - 1, LHY = LTY + LW - 1;  // halo block 42 x 26
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct LossParams {
    float win[LW];         // the 11 normalised Gaussian taps, by value (kernel parameters live in the constant bank of
                           // whichever device the launch goes to: no per-device __constant__ symbol to initialise)
    const float* mask;     // [H,W] or NULL: gt and pred are both multiplied by it (freegaussian_model.py:957-963)
    int W, H, rstride;     // rstride: floats per pixel of `render` (3 or 4 ...)
    const float* render;   // [H,W,rstride]
    const float* alpha;    // [H,W]
    const float* bg;       // [3]
    const float* gt;       // [H,W,3]
    float l1_scale;        // (1 - lambda) / (3 H W)
    float ssim_scale;      // lambda / (3 (H-10) (W-10))
    float* partial;        // [3 maps][3 ch][H-10][W-10]
    double* sums;          // [2]: L1 term, SSIM term (already scaled)
    const float* v_loss;   // scalar upstream gradient (device)
    float* v_render;       // [H,W,rstride]
    float* v_alpha;        // [H,W]
};

__device__ __forceinline__ float pred_at(const LossParams& p, int x, int y, int ch, float& raw) {
    const size_t pix = (size_t)y * p.W + x;
    raw = p.render[pix * p.rstride + ch] + (1.f - p.alpha[pix]) * p.bg[ch];
    return fminf(fmaxf(raw, 0.f), 1.f);
}
__device__ __forceinline__ float mask_at(const LossParams& p, int x, int y) {
    return p.mask ? p.mask[(size_t)y * p.W + x] : 1.f;
}

__global__ void __launch_bounds__(256) l1_ssim_fwd_kernel(LossParams p) {
    pdl_wait();
    __shared__ float sx[LHY][LHX + 1], sy[LHY][LHX + 1];
    __shared__ float hm[5][LHY][LTX + 1];  // horizontally filtered moments
    __shared__ double red[2][8];
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * LTX, y0 = blockIdx.y * LTY;
    const int tid = threadIdx.x;
    const int VW = p.W - (LW - 1), VH = p.H - (LW - 1);
    float l1 = 0.f;
    // halo load: pixel (x0+i, y0+j), i<42, j<26 (zero outside the image: never used by valid outputs)
    for (int q = tid; q < LHX * LHY; q += 256) {
        const int j = q / LHX, i = q - j * LHX;
        const int x = x0 + i, y = y0 + j;
        float xv = 0.f, yv = 0.f;
        if (x < p.W && y < p.H) {
            float raw;
            const float mk = mask_at(p, x, y);
            yv = pred_at(p, x, y, ch, raw) * mk;
            xv = p.gt[((size_t)y * p.W + x) * 3 + ch] * mk;
            if (i < LTX && j < LTY) l1 += fabsf(xv - yv);  // every pixel is the interior of exactly one block
        }
        sx[j][i] = xv;
        sy[j][i] = yv;
    }
    __syncthreads();
    for (int q = tid; q < LTX * LHY; q += 256) {
        const int j = q / LTX, i = q - j * LTX;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < LW; ++k) {
            const float w = p.win[k], xv = sx[j][i + k], yv = sy[j][i + k];
            a = fmaf(w, xv, a); b = fmaf(w, yv, b);
            aa = fmaf(w * xv, xv, aa); bb = fmaf(w * yv, yv, bb); ab = fmaf(w * xv, yv, ab);
        }
        hm[0][j][i] = a; hm[1][j][i] = b; hm[2][j][i] = aa; hm[3][j][i] = bb; hm[4][j][i] = ab;
    }
    __syncthreads();
    float ssim_sum = 0.f;
    for (int q = tid; q < LTX * LTY; q += 256) {
        const int j = q / LTX, i = q - j * LTX;
        const int ox = x0 + i, oy = y0 + j;
        if (ox >= VW || oy >= VH) continue;
        float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
        for (int k = 0; k < LW; ++k) {
            const float w = p.win[k];
            mu1 = fmaf(w, hm[0][j + k][i], mu1); mu2 = fmaf(w, hm[1][j + k][i], mu2);
            exx = fmaf(w, hm[2][j + k][i], exx); eyy = fmaf(w, hm[3][j + k][i], eyy);
            exy = fmaf(w, hm[4][j + k][i], exy);
        }
        // X = gt (constant), Y = pred.  S = (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))
        const float s1 = exx - mu1 * mu1, s2 = eyy - mu2 * mu2, s12 = exy - mu1 * mu2;
        const float n1 = 2.f * mu1 * mu2 + SSIM_C1, n2 = 2.f * s12 + SSIM_C2;
        const float d1 = mu1 * mu1 + mu2 * mu2 + SSIM_C1, d2 = s1 + s2 + SSIM_C2;
        const float S = (n1 * n2) / (d1 * d2);
        ssim_sum += S;
        // partials w.r.t. the moments of Y (pred): mu2, E[yy], E[xy]
        //   dS/dn1 = S/n1, dS/dn2 = S/n2, dS/dd1 = -S/d1, dS/dd2 = -S/d2
        //   n1: d/dmu2 = 2 mu1;  n2: d/dexy = 2, d/dmu2 = -2 mu1;  d1: d/dmu2 = 2 mu2;  d2: d/deyy = 1, d/dmu2 = -2 mu2
        const float inv_d = 1.f / (d1 * d2);
        const float dS_dn1 = n2 * inv_d, dS_dn2 = n1 * inv_d, dS_dd1 = -S / d1, dS_dd2 = -S / d2;
        const float g_mu2 = dS_dn1 * 2.f * mu1 + dS_dn2 * (-2.f * mu1) + dS_dd1 * 2.f * mu2 + dS_dd2 * (-2.f * mu2);
        const float g_eyy = dS_dd2;
        const float g_exy = dS_dn2 * 2.f;
        // loss = ... + lambda * (1 - mean S): fold the -lambda / count in here
        const size_t o = ((size_t)ch * VH + oy) * VW + ox;
        const size_t plane = (size_t)3 * VH * VW;
        p.partial[o] = -p.ssim_scale * g_mu2;
        p.partial[plane + o] = -p.ssim_scale * g_eyy;
        p.partial[2 * plane + o] = -p.ssim_scale * g_exy;
    }
    // block reduction of both sums (double accumulators keep the scalar loss reproducible to ~1e-7)
    double v0 = (double)l1 * p.l1_scale, v1 = (double)ssim_sum * p.ssim_scale;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = v0; red[1][tid >> 5] = v1; }
    __syncthreads();
    if (tid == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(p.sums, a);
        atomicAdd(p.sums + 1, b);
    }
}

__global__ void __launch_bounds__(256) l1_ssim_bwd_kernel(LossParams p) {
    pdl_wait();
    // input pixel block 32x16; needs partial maps at q in [p-10, p] -> halo to the top/left
    __shared__ float sg[3][LHY][LHX + 1];
    __shared__ float hg[3][LHY][LTX + 1];
    const int x0 = blockIdx.x * LTX, y0 = blockIdx.y * LTY;
    const int tid = threadIdx.x;
    const int VW = p.W - (LW - 1), VH = p.H - (LW - 1);
    const size_t plane = (size_t)3 * VH * VW;
    const float vl = *p.v_loss;
    float acc_alpha[2] = {0.f, 0.f};  // each thread owns 2 pixels: q = tid, tid + 256
    for (int ch = 0; ch < 3; ++ch) {
        __syncthreads();
        for (int q = tid; q < LHX * LHY; q += 256) {
            const int j = q / LHX, i = q - j * LHX;
            const int qx = x0 + i - (LW - 1), qy = y0 + j - (LW - 1);
            const bool ok = qx >= 0 && qy >= 0 && qx < VW && qy < VH;
            const size_t o = ((size_t)ch * VH + (ok ? qy : 0)) * VW + (ok ? qx : 0);
#pragma unroll
            for (int m = 0; m < 3; ++m) sg[m][j][i] = ok ? p.partial[m * plane + o] : 0.f;
        }
        __syncthreads();
        // horizontal: out(i) = sum_k w[k] g(i + (LW-1) - k)   (pixel x gets q = x - k)
        for (int q = tid; q < LTX * LHY; q += 256) {
            const int j = q / LTX, i = q - j * LTX;
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < LW; ++k) a = fmaf(p.win[k], sg[m][j][i + (LW - 1) - k], a);
                hg[m][j][i] = a;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = tid + r * 256;
            const int j = q / LTX, i = q - j * LTX;
            const int x = x0 + i, y = y0 + j;
            if (x >= p.W || y >= p.H) continue;
            float t[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < LW; ++k) a = fmaf(p.win[k], hg[m][j + (LW - 1) - k][i], a);
                t[m] = a;
            }
            float raw;
            const float mk = mask_at(p, x, y);
            const float pr = pred_at(p, x, y, ch, raw) * mk;
            const float g = p.gt[((size_t)y * p.W + x) * 3 + ch] * mk;
            // d loss / d pred: SSIM part through mu2, E[yy] (2 pred), E[xy] (gt) + L1 part
            float d = t[0] + 2.f * pr * t[1] + g * t[2];
            d += p.l1_scale * ((pr > g) ? 1.f : ((pr < g) ? -1.f : 0.f));
            d *= vl * mk;
            if (!(raw >= 0.f && raw <= 1.f)) d = 0.f;  // clamp backward
            const size_t pix = (size_t)y * p.W + x;
            p.v_render[pix * p.rstride + ch] = d;
            acc_alpha[r] -= d * p.bg[ch];
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = tid + r * 256;
        const int j = q / LTX, i = q - j * LTX;
        const int x = x0 + i, y = y0 + j;
        if (x >= p.W || y >= p.H) continue;
        const size_t pix = (size_t)y * p.W + x;
        p.v_alpha[pix] = acc_alpha[r];
        for (int c = 3; c < p.rstride; ++c) p.v_render[pix * p.rstride + c] = 0.f;
    }
}

static void set_window(LossParams& p) {
    float w[LW];
    for (int i = 0; i < LW; ++i) { double c = i - LR; w[i] = (float)exp(-(c * c) / (2.0 * 1.5 * 1.5)); }
    // normalise in float like torch: g / g.sum()
    float fs = 0.f;
    for (int i = 0; i < LW; ++i) fs += w[i];
    for (int i = 0; i < LW; ++i) p.win[i] = w[i] / fs;
}

// ---- depth fix-up of freegaussian_model.py:884-886: depth = where(alpha > 0, ED, max over the image of ED) ----
__device__ __forceinline__ uint32_t float_to_ordered(float f) {  // order-preserving map float -> uint32
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void __launch_bounds__(256) depth_max_kernel(long long n, const float* render, int stride, int ch, uint32_t* out) {
    pdl_wait();
    uint32_t m = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        m = max(m, float_to_ordered(render[i * stride + ch]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
__global__ void __launch_bounds__(256) depth_fixup_kernel(long long n, const float* render, int stride, int ch,
                                                          const float* alpha, const uint32_t* mx, float* depth) {
    pdl_wait();
    const float big = ordered_to_float(*mx);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        depth[i] = alpha[i] > 0.f ? render[i * stride + ch] : big;
}
__global__ void __launch_bounds__(256) depth_fixup_bwd_kernel(long long n, const float* alpha, const float* v_depth,
                                                              int stride, int ch, float* v_render) {
    pdl_wait();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        for (int c = 0; c < stride; ++c) v_render[i * stride + c] = (c == ch && alpha[i] > 0.f) ? v_depth[i] : 0.f;
    }
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_l1_ssim_workspace_floats(int width, int height) {
    if (width < LW || height < LW) return 0;
    return (int64_t)9 * (width - (LW - 1)) * (height - (LW - 1));
}

extern "C" int fg_l1_ssim_fwd(int width, int height, int render_stride, const float* render, const float* alpha,
                              const float* background, const float* gt, const float* mask, float ssim_lambda,
                              float* partial, double* sums /*[2], zeroed inside*/, void* stream) {
    FG_REQUIRE(width >= LW && height >= LW, "image smaller than the 11x11 SSIM window");
    FG_REQUIRE(render_stride >= 3 && render && alpha && background && gt && partial && sums, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    FG_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    LossParams p = {};
    set_window(p);
    p.mask = mask;
    p.W = width; p.H = height; p.rstride = render_stride; p.render = render; p.alpha = alpha; p.bg = background; p.gt = gt;
    p.l1_scale = (1.f - ssim_lambda) / (3.f * width * height);
    p.ssim_scale = ssim_lambda / (3.f * (float)(width - (LW - 1)) * (float)(height - (LW - 1)));
    p.partial = partial; p.sums = sums;
    dim3 grid((width + LTX - 1) / LTX, (height + LTY - 1) / LTY, 3);
    FG_LAUNCH(l1_ssim_fwd_kernel, grid, 256, 0, st, p);
    return FG_OK;
}

extern "C" int fg_l1_ssim_bwd(int width, int height, int render_stride, const float* render, const float* alpha,
                              const float* background, const float* gt, const float* mask, float ssim_lambda,
                              const float* partial, const float* v_loss, float* v_render, float* v_alpha, void* stream) {
    FG_REQUIRE(width >= LW && height >= LW, "image smaller than the 11x11 SSIM window");
    FG_REQUIRE(render_stride >= 3 && render && alpha && background && gt && partial && v_loss && v_render && v_alpha,
               "bad arguments");
    LossParams p = {};
    set_window(p);
    p.mask = mask;
    p.W = width; p.H = height; p.rstride = render_stride; p.render = render; p.alpha = alpha; p.bg = background; p.gt = gt;
    p.l1_scale = (1.f - ssim_lambda) / (3.f * width * height);
    p.ssim_scale = ssim_lambda / (3.f * (float)(width - (LW - 1)) * (float)(height - (LW - 1)));
    p.partial = const_cast<float*>(partial); p.v_loss = v_loss; p.v_render = v_render; p.v_alpha = v_alpha;
    dim3 grid((width + LTX - 1) / LTX, (height + LTY - 1) / LTY, 1);
    FG_LAUNCH(l1_ssim_bwd_kernel, grid, 256, 0, stream, p);
    return FG_OK;
}

extern "C" int fg_depth_fixup_fwd(int64_t n_pixels, const float* render, int render_stride, int channel, const float* alpha,
                                  float* depth, uint32_t* max_ws /*[1]*/, void* stream) {
    FG_REQUIRE(n_pixels >= 0 && render_stride >= 1 && channel >= 0 && channel < render_stride, "bad shape");
    if (n_pixels == 0) return FG_OK;
    FG_REQUIRE(render && alpha && depth && max_ws, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    FG_CUDA(cudaMemsetAsync(max_ws, 0, 4, st));
    const int grid = (int)std::min<long long>((n_pixels + 255) / 256, num_sms() * 8);
    FG_LAUNCH(depth_max_kernel, grid, 256, 0, st, (long long)n_pixels, render, render_stride, channel, max_ws);
    FG_LAUNCH(depth_fixup_kernel, grid, 256, 0, st, (long long)n_pixels, render, render_stride, channel, alpha, max_ws, depth);
    return FG_OK;
}

extern "C" int fg_depth_fixup_bwd(int64_t n_pixels, const float* alpha, const float* v_depth, int render_stride,
                                  int channel, float* v_render, void* stream) {
    FG_REQUIRE(n_pixels >= 0 && render_stride >= 1 && channel >= 0 && channel < render_stride, "bad shape");
    if (n_pixels == 0) return FG_OK;
    FG_REQUIRE(alpha && v_depth && v_render, "NULL pointer");
    const int grid = (int)std::min<long long>((n_pixels + 255) / 256, num_sms() * 8);
    FG_LAUNCH(depth_fixup_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, (long long)n_pixels, alpha, v_depth, render_stride,
              channel, v_render);
    return FG_OK;
}
