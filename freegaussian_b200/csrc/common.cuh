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

// SM count of the current device (148 on a B200), queried once per device: grids of the persistent / grid-stride
// kernels are sized from it.
int num_sms();

// run-time switches (fg_set_option, api.cu)
extern int g_fwd_two_pixels;   // rasterize_fwd.cu: 1 = two-pixel packed forward kernel
extern int g_xchg_ar_blocks;    // exchange.cu: CTAs of the all-reduce kernel
extern int g_xchg_pull_blocks;  // exchange.cu: CTAs per peer of the pull kernel

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// first statement of every kernel (see FG_LAUNCH); a no-op under a plain launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
// Every launch is a programmatic dependent launch: the kernel may be scheduled while the previous kernel
// of the stream drains (saves ~2 us of launch latency per kernel, ~50 launches per step), and therefore
// EVERY kernel of this library starts with fg::pdl_wait() before touching memory.
#define FG_LAUNCH(kernel, grid, block, smem, strm__, ...)                                      \
    do {                                                                                       \
        cudaLaunchConfig_t cfg__ = {};                                                         \
        cfg__.gridDim = dim3(grid);                                                            \
        cfg__.blockDim = dim3(block);                                                          \
        cfg__.dynamicSmemBytes = (smem);                                                       \
        cfg__.stream = (cudaStream_t)(strm__);                                                 \
        cudaLaunchAttribute at__[1];                                                           \
        at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                       \
        at__[0].val.programmaticStreamSerializationAllowed = 1;                                \
        cfg__.attrs = at__;                                                                    \
        cfg__.numAttrs = 1;                                                                    \
        cudaError_t e__ = cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                     \
        fg::g_launch_count.fetch_add(1, std::memory_order_relaxed);                            \
        if (e__ != cudaSuccess) return fg::set_cuda_error(e__, __FILE__, __LINE__);            \
    } while (0)
