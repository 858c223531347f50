// Error reporting, ABI version and launch accounting for the C ABI (include/fg_api.h).
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace fg {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

int set_error(int code, const char* msg, const char* file, int line) {
    const char* base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "%s (%s:%d)", msg, base ? base + 1 : file, line);
    return code;
}

int set_cuda_error(cudaError_t e, const char* file, int line) {
    const char* base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "CUDA error %d: %s (%s:%d)", (int)e, cudaGetErrorString(e),
             base ? base + 1 : file, line);
    return FG_ERR_CUDA;
}

}  // namespace fg

namespace fg {
int g_xchg_ar_blocks = 64;  // 8 ranks, 68 MB: 16 / 32 / 64 CTAs -> 0.256 / 0.261 / 0.245 ms beside the SH-row kernels
int g_xchg_pull_blocks = 4;
}  // namespace fg

/* Run-time switches for A/B measurements (bench.py, tests). */
extern "C" int fg_set_option(const char* name, int value) {
    FG_REQUIRE(name != nullptr, "name must not be NULL");
    if (strcmp(name, "fwd_two_pixels") == 0) { fg::g_fwd_two_pixels = value != 0; return FG_OK; }
    if (strcmp(name, "xchg_ar_blocks") == 0) { FG_REQUIRE(value >= 1 && value <= 255, "1..255"); fg::g_xchg_ar_blocks = value; return FG_OK; }
    if (strcmp(name, "xchg_pull_blocks") == 0) { FG_REQUIRE(value >= 1 && value <= 64, "1..64"); fg::g_xchg_pull_blocks = value; return FG_OK; }
    FG_REQUIRE(false, "unknown option");
}

extern "C" {

const char* fg_last_error(void) { return fg::g_err; }
int fg_abi_version(void) { return 12; }
long long fg_launch_count(void) { return fg::g_launch_count.load(); }

}  // extern "C"

// ---------------------------------------------------------------------------------------
// FP32 FMA peak microbenchmark: the denominator for the FP32-pipe-bound compositing kernels
// (MEASURED_PEAKS.json has only the HBM copy and the bf16 GEMM figures).  8 independent FFMA
// chains per thread, 1024 threads per SM-resident CTA, 2 CTAs per SM.
namespace fg {
__global__ void __launch_bounds__(1024) fp32_peak_kernel(float* out, int iters, float a, float b) {
    pdl_wait();
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s;  // keep the chains alive
}
}  // namespace fg

extern "C" int fg_measure_fp32_tflops(double* tflops_host, void* stream) {
    FG_REQUIRE(tflops_host, "tflops_host must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    float* d = nullptr;
    FG_CUDA(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    FG_CUDA(cudaEventCreate(&e0));
    FG_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = fg::num_sms() * 2;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        FG_CUDA(cudaEventRecord(e0, st));
        FG_LAUNCH(fg::fp32_peak_kernel, blocks, 1024, 0, st, d, iters, 1.0000001f, 1e-9f);
        FG_CUDA(cudaEventRecord(e1, st));
        FG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        FG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * iters * 1024.0 * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_host = best;
    return FG_OK;
}
