// Shared plumbing for the sm_100a kernels: error reporting across the C ABI, launch
// accounting (bench.py's `gpu_launches`), small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/fg_api.h"

namespace fg {

int set_error(int code, const char* msg, const char* file, int line);
int set_cuda_error(cudaError_t e, const char* file, int line);
extern std::atomic<long long> g_launch_count;

constexpr int kNumSMs = 148;  // B200

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace fg

#define FG_REQUIRE(cond, msg)                                                       \
    do {                                                                            \
        if (!(cond)) return fg::set_error(FG_ERR_INVALID, msg, __FILE__, __LINE__); \
    } while (0)

#define FG_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return fg::set_cuda_error(e__, __FILE__, __LINE__); \
    } while (0)

// Launch + count + check.  All kernels go through this so gpu_launches is exact.
#define FG_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                          \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
        fg::g_launch_count.fetch_add(1, std::memory_order_relaxed);               \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) return fg::set_cuda_error(e__, __FILE__, __LINE__); \
    } while (0)
