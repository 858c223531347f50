// Error reporting, ABI version and launch accounting for the C ABI (include/fg_api.h).
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace fg {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

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

extern "C" {

const char* fg_last_error(void) { return fg::g_err; }
int fg_abi_version(void) { return 1; }
long long fg_launch_count(void) { return fg::g_launch_count.load(); }

}  // extern "C"
