// Attribute-mask assignment of preprocess/knn_gaussian.py:116-132: after a
// packed "ED" render of a key frame, every visible Gaussian whose projected centre lands on a pixel
// with a consistent depth inherits that pixel's attribute masks.  One thread per visible splat.
//
//   xy   = trunc(means2d)                      (torch .long(): truncation toward zero)        :116
//   keep = 0 <= xy < (W,H)                                                                      :117
//   D    = depth[xy.y, xy.x];  delta = D - z;  keep &= (-0.1 D < delta) & (delta < D)          :120-122
//   gaussian_masks[gaussian_id, m] = True  for every attribute m with mask[xy.y, xy.x, m]      :127-132
//
// Roofline: HBM, 24 B read per visible splat + M mask bytes; the writes are sparse byte stores
// (all writers store the same value, so no atomics are needed).
#include "common.cuh"

namespace fg {

__global__ void __launch_bounds__(256)
    assign_masks_kernel(long long nnz, const float2* __restrict__ means2d, const float* __restrict__ depths,
                        const int64_t* __restrict__ gaussian_ids, const float* __restrict__ depth_img, int W, int H,
                        const uint8_t* __restrict__ atrb_masks, const uint8_t* __restrict__ mask_valids, int M,
                        uint8_t* __restrict__ gaussian_masks) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= nnz) return;
    const float2 m = means2d[i];
    const long long x = (long long)m.x, y = (long long)m.y;  // truncation toward zero, like tensor.long()
    if (x < 0 || x >= W || y < 0 || y >= H) return;
    const float D = depth_img[y * W + x];
    const float delta = D - depths[i];
    if (!((-D * 0.1f < delta) && (delta < D * 1.f))) return;
    const uint8_t* px = atrb_masks + ((size_t)y * W + x) * M;
    uint8_t* row = gaussian_masks + (size_t)gaussian_ids[i] * M;
    for (int a = 0; a < M; ++a)
        if (px[a] && mask_valids[a]) row[a] = 1;
}

}  // namespace fg

extern "C" int fg_assign_masks(int64_t nnz, const float* means2d, const float* depths, const int64_t* gaussian_ids,
                               const float* depth_img, int width, int height, const uint8_t* atrb_masks,
                               const uint8_t* mask_valids, int n_attr, uint8_t* gaussian_masks, void* stream) {
    FG_REQUIRE(nnz >= 0 && width > 0 && height > 0 && n_attr >= 0, "bad sizes");
    if (nnz == 0 || n_attr == 0) return FG_OK;
    FG_REQUIRE(means2d && depths && gaussian_ids && depth_img && atrb_masks && mask_valids && gaussian_masks, "NULL pointer");
    FG_LAUNCH(fg::assign_masks_kernel, fg::ceil_div(nnz, 256), 256, 0, stream, (long long)nnz, (const float2*)means2d,
              depths, gaussian_ids, depth_img, width, height, atrb_masks, mask_valids, n_attr, gaussian_masks);
    return FG_OK;
}
