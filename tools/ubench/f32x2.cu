// Microbenchmark: does packed FP32 (fma.rn.f32x2 -> FFMA2) free issue slots on sm_100a?
//   A: 8 independent FFMA chains per thread                      (flops = 2 per instr per lane)
//   B: 8 independent FFMA2 chains per thread                     (flops = 4 per instr per lane)
//   C: A interleaved 1:1 with integer ALU ops (LOP3/IADD3)        (how much does issue contention cost?)
//   D: B interleaved with the same number of ALU ops per FLOP as C
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float a, float b, unsigned m) {
    float x[8];
    float2 y[8];
    unsigned z[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = make_float2(x[i], x[i] + 0.5f); z[i] = threadIdx.x * 7 + i; }
    const float2 a2 = make_float2(a, a * 1.0000001f), b2 = make_float2(b, b * 2.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 2) x[i] = fmaf(x[i], a, b);
                if (MODE == 1 || MODE == 3) y[i] = __ffma2_rn(y[i], a2, b2);
                if (MODE == 2) z[i] = (z[i] ^ m) + (z[i] >> 3);
                if (MODE == 3) { z[i] = (z[i] ^ m) + (z[i] >> 3); z[i] = (z[i] ^ (m >> 1)) + (z[i] >> 5); }
            }
        }
    }
    float s = 0; unsigned t = 0;
    for (int i = 0; i < 8; ++i) { s += x[i] + y[i].x + y[i].y; t += z[i]; }
    if (s == 123.456f || t == 0x12345u) out[0] = s + t;
}

template <int MODE>
void run(const char* name, double flop_per_iter_per_thread) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev, sms; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 2048, blocks = sms * 2;
    double best = 1e30;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 1024>>>(d, iters, 1.0000001f, 1e-9f, 0x5bd1e995u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    double flops = flop_per_iter_per_thread * iters * 1024.0 * blocks;
    printf("%-40s %8.3f ms  %7.2f TFLOP/s\n", name, best, flops / (best * 1e-3) / 1e12);
    cudaFree(d);
}

int main() {
    run<0>("A  FFMA x64 / iter", 2.0 * 64);
    run<1>("B  FFMA2 x64 / iter", 4.0 * 64);
    run<2>("C  FFMA x64 + ~128 ALU / iter", 2.0 * 64);
    run<3>("D  FFMA2 x64 + ~256 ALU / iter", 4.0 * 64);
    return 0;
}
