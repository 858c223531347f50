// packed=True layout of the render call (callers: preprocess/knn_gaussian.py:93-113,
// render_depth.py:99-118, render_color.py:93-112, o3d_color_splat.py:188-208): every per-(camera, Gaussian) tensor is
// compacted to the visible pairs in ascending c*N+n order and the tile lists index the compacted rows.  Three launches
// (flags+scan, gather, remap) instead of nonzero / cumsum / seven fancy-index gathers in torch.
// Roofline: HBM, ~100 B per visible pair + 8 B per tile intersection.
#include <algorithm>

#include "common.cuh"

extern "C" int fg_exclusive_scan_i32(int64_t n, const int32_t* counts, int32_t* offsets, int64_t* total,
                                     void* workspace, int64_t workspace_bytes, void* stream);
extern "C" int64_t fg_scan_workspace_bytes(int64_t n);

namespace fg {

__global__ void __launch_bounds__(256) pack_flag_kernel(long long total, const int32_t* __restrict__ radii, int32_t* __restrict__ flags) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < total) flags[i] = radii[i] > 0 ? 1 : 0;
}

struct PackArgs {
    int N, CH, opac_shared;
    const int32_t* radii;
    const int32_t* offsets;
    const float2* means2d;
    const float* depths;
    const float* conics;
    const float* feat;
    const float* opac;
    const float4* aff;
    int32_t* radii_p;
    float2* means2d_p;
    float* depths_p;
    float* conics_p;
    float* feat_p;
    float* opac_p;
    float4* aff_p;
    long long* camera_ids;
    long long* gaussian_ids;
};

__global__ void __launch_bounds__(256) pack_gather_kernel(long long total, PackArgs a) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int r = a.radii[i];
    if (r <= 0) return;
    const int s = a.offsets[i];
    const long long c = i / a.N, n = i - c * a.N;
    a.radii_p[s] = r;
    a.means2d_p[s] = a.means2d[i];
    a.depths_p[s] = a.depths[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) a.conics_p[3 * (size_t)s + k] = a.conics[3 * (size_t)i + k];
    for (int k = 0; k < a.CH; ++k) a.feat_p[(size_t)s * a.CH + k] = a.feat[(size_t)i * a.CH + k];
    a.opac_p[s] = a.opac[a.opac_shared ? n : i];
    if (a.aff) a.aff_p[s] = a.aff[i];
    a.camera_ids[s] = c;
    a.gaussian_ids[s] = n;
}

__global__ void __launch_bounds__(256) pack_remap_kernel(long long M, const int32_t* __restrict__ offsets, int32_t* __restrict__ ids) {
    pdl_wait();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < M; i += (long long)gridDim.x * 256) ids[i] = offsets[ids[i]];
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_pack_workspace_bytes(int64_t total) {
    const int64_t n = total < 1 ? 1 : total;
    return ((n * 4 + 255) & ~(int64_t)255) + fg_scan_workspace_bytes(n);
}

extern "C" int fg_pack_plan(int64_t total, const int32_t* radii, int32_t* offsets, int64_t* nnz_dev, void* workspace,
                            int64_t workspace_bytes, void* stream) {
    FG_REQUIRE(total >= 0 && total < (1ll << 31) && nnz_dev, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (total == 0) {
        FG_CUDA(cudaMemsetAsync(nnz_dev, 0, 8, st));
        return FG_OK;
    }
    FG_REQUIRE(radii && offsets && workspace && workspace_bytes >= fg_pack_workspace_bytes(total), "NULL pointer / workspace too small");
    int32_t* flags = (int32_t*)workspace;
    unsigned char* scan_ws = (unsigned char*)workspace + (((size_t)total * 4 + 255) & ~(size_t)255);
    FG_LAUNCH(pack_flag_kernel, ceil_div(total, 256), 256, 0, st, (long long)total, radii, flags);
    return fg_exclusive_scan_i32(total, flags, offsets, nnz_dev, scan_ws, fg_scan_workspace_bytes(total), stream);
}

extern "C" int fg_pack_gather(int C, int N, int CH, const int32_t* radii, const int32_t* offsets, const float* means2d,
                              const float* depths, const float* conics, const float* feat, const float* opacities,
                              int opac_shared, const float* flow_affine, int32_t* radii_p, float* means2d_p, float* depths_p,
                              float* conics_p, float* feat_p, float* opac_p, float* flow_affine_p, int64_t* camera_ids,
                              int64_t* gaussian_ids, void* stream) {
    FG_REQUIRE(C >= 1 && N >= 0 && CH >= 0, "bad C / N / CH");
    const long long total = (long long)C * N;
    if (total == 0) return FG_OK;
    FG_REQUIRE(radii && offsets && means2d && depths && conics && opacities && (feat || CH == 0), "NULL input pointer");
    FG_REQUIRE(!flow_affine || flow_affine_p, "flow_affine_p must mirror flow_affine");
    PackArgs a;
    a.N = N; a.CH = CH; a.opac_shared = opac_shared; a.radii = radii; a.offsets = offsets; a.means2d = (const float2*)means2d;
    a.depths = depths; a.conics = conics; a.feat = feat; a.opac = opacities; a.aff = (const float4*)flow_affine;
    a.radii_p = radii_p; a.means2d_p = (float2*)means2d_p; a.depths_p = depths_p; a.conics_p = conics_p; a.feat_p = feat_p;
    a.opac_p = opac_p; a.aff_p = (float4*)flow_affine_p; a.camera_ids = (long long*)camera_ids; a.gaussian_ids = (long long*)gaussian_ids;
    FG_LAUNCH(pack_gather_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, total, a);
    return FG_OK;
}

extern "C" int fg_pack_remap(int64_t M, const int32_t* offsets, int32_t* flatten_ids, void* stream) {
    FG_REQUIRE(M >= 0, "bad M");
    if (M == 0) return FG_OK;
    FG_REQUIRE(offsets && flatten_ids, "NULL pointer");
    const int grid = (int)std::min<long long>((M + 255) / 256, (long long)num_sms() * 16);
    FG_LAUNCH(pack_remap_kernel, grid, 256, 0, (cudaStream_t)stream, (long long)M, offsets, flatten_ids);
    return FG_OK;
}
