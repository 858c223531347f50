// Host-side orchestration of the forward pass in two C calls (the native runtime of the path):
//   fg_render_front: projection -> depth sort -> bin count -> tile scan -> coarse (cell) scan -> the ONE host
//                    sync (M = tile intersections, Mc = coarse pairs size the list buffers)
//   fg_render_back : coarse pairs by cell (ranked placement, or emit -> sort -> cell offsets when there are more than
//                    1024 coarse cells) -> fine binning -> compositing forward
// Same kernels as the granular entry points (which stay for tests, the other list-building modes
// and per-stage timing); what this removes is ~10 Python/ctypes round trips and a dozen temporary
// allocations per step: on the 10 k-Gaussian cfg1 scene a step is host-bound (1.0 ms against
// 0.38 ms of kernels), and after every host sync the GPU waits for the host to catch up.
#include "common.cuh"

namespace fg {

static inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct FrontLayout {
    size_t diff, ccnt, n2, scan, tscan, sort, rank, coff, total;
    size_t rank_bytes;  // > 0: the ranked placement of binning.cu applies (few enough coarse cells)
};
static FrontLayout front_layout(int C, int N, int tile_w, int tile_h) {
    const size_t total = (size_t)C * N;
    FrontLayout L;
    size_t o = 0;
    L.diff = o; o += al((size_t)C * (tile_h + 1) * (tile_w + 1) * 4);
    L.ccnt = o; o += al(total * 4);
    L.n2 = o; o += al(32);
    L.scan = o; o += al((size_t)fg_scan_workspace_bytes((int64_t)total));
    L.tscan = o; o += al((size_t)fg_bin_tile_scan_workspace_bytes(C, tile_w, tile_h));
    L.sort = o; o += al((size_t)fg_depth_sort_workspace_bytes((int64_t)total));
    // read by fg_render_back: the (chunk, cell) prefix matrix and the cell offsets of the ranked placement
    L.rank_bytes = (size_t)fg_bin_ranked_workspace_bytes(C, N, tile_w, tile_h);
    L.rank = o; o += al(L.rank_bytes);
    int cw = 0, chh = 0;
    fg_bin_coarse_dims(tile_w, tile_h, &cw, &chh);
    L.coff = o; o += al(L.rank_bytes ? ((size_t)C * cw * chh + 1) * 4 : 0);
    L.total = o;
        return L;
}

struct BackLayout {
    size_t ck, cv, ck2, cv2, coff, sort, total;
};
static BackLayout back_layout(int C, int tile_w, int tile_h, int64_t Mc) {
    int cw = 0, chh = 0;
    fg_bin_coarse_dims(tile_w, tile_h, &cw, &chh);
    const size_t m = (size_t)(Mc > 0 ? Mc : 1);
    BackLayout L;
    size_t o = 0;
    L.ck = o; o += al(m * 4);
    L.cv = o; o += al(m * 4);
    L.ck2 = o; o += al(m * 4);
    L.cv2 = o; o += al(m * 4);
    L.coff = o; o += al((size_t)C * cw * chh * 4);
    L.sort = o; o += al((size_t)fg_radix_sort_workspace_bytes((int64_t)m));
    L.total = o;
        return L;
}

// One pinned, device-mapped slot per host thread (portable: valid under every device's context), so two threads / devices
// rendering concurrently never share one: {M, Mc, sequence number} written by the device, read by the host.
struct CountSlot {
    volatile int64_t v[4];
};
static thread_local int64_t g_count_seq = 0;  // calls made through this thread's slot
static CountSlot* count_slot() {
    static thread_local CountSlot* p = nullptr;
    if (!p) {
        if (cudaHostAlloc((void**)&p, sizeof(CountSlot), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) return p = nullptr;
        p->v[0] = p->v[1] = p->v[2] = p->v[3] = 0;
    }
    return p;
}

// The list sizes reach the host through a store into mapped host memory, and the host polls for the sequence number:
// the GPU idles between the scans and the list building for as long as the host needs to learn M, and a
// cudaMemcpyAsync + cudaStreamSynchronize round trip took 20 us alone and 75 us with an input prefetch (H2D copy) in flight
// on another stream (bench.py e2e, round 2); a posted write + a polling load takes ~3 us either way.
__global__ void publish_counts_kernel(const int64_t* __restrict__ n2, volatile int64_t* slot, int64_t seq) {
    pdl_wait();
    slot[0] = n2[0];
    slot[1] = n2[1];
    __threadfence_system();
    slot[2] = seq;
}

static int publish_counts_launch(const int64_t* n2, CountSlot* slot, int64_t seq, cudaStream_t st) {
    int64_t* dev_slot = nullptr;
    FG_CUDA(cudaHostGetDevicePointer((void**)&dev_slot, (void*)slot, 0));
    FG_LAUNCH(publish_counts_kernel, 1, 1, 0, st, n2, (volatile int64_t*)dev_slot, seq);
    return FG_OK;
}

// wait until the kernel above has run: poll the slot; every few thousand polls ask the driver whether the stream died
static int wait_counts(CountSlot* slot, int64_t seq, cudaStream_t st) {
    for (unsigned spins = 1;; ++spins) {
        if (slot->v[2] == seq) return FG_OK;
        if ((spins & 0x3fff) == 0) {
            const cudaError_t q = cudaStreamQuery(st);
            if (q == cudaSuccess) {  // everything ran: the store has landed (or never will)
                if (slot->v[2] == seq) return FG_OK;
                FG_CUDA(cudaStreamSynchronize(st));
                FG_REQUIRE(slot->v[2] == seq, "the list sizes never reached the host");
                return FG_OK;
            }
            if (q != cudaErrorNotReady) return set_cuda_error(q, __FILE__, __LINE__);
        }
    }
}

}  // namespace fg

using namespace fg;

extern "C" int64_t fg_render_front_workspace_bytes(int C, int N, int tile_w, int tile_h) {
    return (int64_t)front_layout(C, N, tile_w, tile_h).total;
}

extern "C" int fg_render_front(int C, int N, const float* means, const float* quats, const float* scales,
                               const float* viewmats, const float* Ks, int width, int height, float eps2d,
                               float near_plane, float far_plane, float radius_clip, int tile_size, int sh_degree,
                               int sh_bases, const float* sh_coeffs, const float* means_next, const float* quats_next,
                               const float* scales_next, int flow_cov, int32_t* radii, float* means2d, float* depths,
                               float* conics, float* compensations, float* feat, int feat_stride, int rgb_off,
                               int depth_off, int flow_off, float* flow_affine, int32_t* tiles_per_gauss,
                               int32_t* order, int32_t* isect_offsets, int32_t* coarse_off, int64_t* counts_host,
                               void* workspace, int64_t workspace_bytes, int32_t* flatten_ids,
                               int64_t flatten_capacity, void* back_workspace, int64_t back_workspace_bytes,
                               void* stream) {
    FG_REQUIRE(isect_offsets && counts_host && workspace, "NULL pointer");
    FG_REQUIRE((long long)C * N == 0 || (order && coarse_off), "order / coarse_off must not be NULL");
    const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
    const FrontLayout L = front_layout(C, N, tile_w, tile_h);
    FG_REQUIRE((size_t)workspace_bytes >= L.total, "front workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)C * N;
    int e;
    if ((e = fg_project_fwd(C, N, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                            radius_clip, tile_size, sh_degree, sh_bases, sh_coeffs, means_next, quats_next, scales_next,
                            flow_cov, radii, means2d, depths, conics, compensations, feat, feat_stride, rgb_off,
                            depth_off, flow_off, flow_affine, tiles_per_gauss, stream)))
        return e;
    // depth order of the visible splats (culled ones are dropped by the first radix pass); n2[2] = their number
    int64_t* n_vis = (int64_t*)(ws + L.n2) + 2;
    if ((e = fg_depth_sort_visible(total, depths, tiles_per_gauss, order, n_vis, ws + L.sort, (int64_t)(L.rank - L.sort), stream)))
        return e;
    int64_t* n2 = (int64_t*)(ws + L.n2);
    const bool ranked = L.rank_bytes > 0;
    if (ranked) {  // few coarse cells: the (splat, cell) pairs are placed by rank, never sorted (binning.cu)
        if ((e = fg_bin_count_cells(C, N, order, means2d, radii, tile_size, tile_w, tile_h, (int32_t*)(ws + L.diff),
                                    ws + L.rank, (int64_t)L.rank_bytes, stream)))
            return e;
    } else if ((e = fg_bin_count(C, N, order, means2d, radii, tile_size, tile_w, tile_h, (int32_t*)(ws + L.diff),
                                 (int32_t*)(ws + L.ccnt), stream)))
        return e;
    if ((e = fg_bin_tile_scan(C, tile_w, tile_h, (int32_t*)(ws + L.diff), isect_offsets, n2, ws + L.tscan,
                              (int64_t)(L.sort - L.tscan), stream)))
        return e;
    if (ranked) {
        if ((e = fg_bin_cell_scan(C, N, tile_w, tile_h, n_vis, ws + L.rank, (int64_t)L.rank_bytes, (int32_t*)(ws + L.coff),
                                  n2 + 1, stream)))
            return e;
    } else if ((e = fg_exclusive_scan_i32(total, (const int32_t*)(ws + L.ccnt), coarse_off, n2 + 1, ws + L.scan,
                                          (int64_t)(L.tscan - L.scan), stream)))
        return e;
    CountSlot* slot = count_slot();
    FG_REQUIRE(slot != nullptr, "cudaHostAlloc failed");
    const int64_t seq = ++g_count_seq;
    if ((e = publish_counts_launch(n2, slot, seq, st))) return e;
    if ((e = wait_counts(slot, seq, st))) return e;  // the one host sync of the forward pass
    const int64_t pin[2] = {slot->v[0], slot->v[1]};
    counts_host[0] = pin[0];
    counts_host[1] = pin[1];
    counts_host[2] = 0;
    // the caller guessed the list size: if it fits, keep the GPU busy without a round trip through the host mirror
    if (flatten_ids && back_workspace && pin[0] <= flatten_capacity &&
        fg_render_back_workspace_bytes(C, tile_w, tile_h, pin[1]) <= back_workspace_bytes) {
        if ((e = fg_render_back(C, N, pin[0], pin[1], order, coarse_off, means2d, radii, tile_size, isect_offsets,
                                flatten_ids, back_workspace, back_workspace_bytes, workspace, workspace_bytes, 0, width,
                                height, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 0, -1, 0, nullptr, nullptr, nullptr,
                                nullptr, stream)))
            return e;
        counts_host[2] = 1;
    }
    return FG_OK;
}

extern "C" int64_t fg_render_back_workspace_bytes(int C, int tile_w, int tile_h, int64_t n_coarse) {
    return (int64_t)back_layout(C, tile_w, tile_h, n_coarse).total;
}

extern "C" int fg_render_back(int C, int N, int64_t n_isects, int64_t n_coarse, const int32_t* order,
                              const int32_t* coarse_off, const float* means2d, const int32_t* radii, int tile_size,
                              const int32_t* isect_offsets, int32_t* flatten_ids, void* workspace,
                              int64_t workspace_bytes, const void* front_workspace, int64_t front_workspace_bytes, int CH,
                              int width, int height, const float* conics,
                              const float* feat, const float* opacities, const float* backgrounds,
                              const float* flow_affine, int flow_ch0, int split, int ed_channel, int opac_shared,
                              float* render, float* render2, float* alphas, int32_t* last_ids, void* stream) {
    FG_REQUIRE(workspace, "workspace must not be NULL");
    const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
    const BackLayout L = back_layout(C, tile_w, tile_h, n_coarse);
    FG_REQUIRE((size_t)workspace_bytes >= L.total, "back workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    int e;
    if (n_isects > 0) {
        FG_REQUIRE(flatten_ids && order && coarse_off, "NULL pointer");
        int cw = 0, chh = 0;
        fg_bin_coarse_dims(tile_w, tile_h, &cw, &chh);
        uint32_t* ck = (uint32_t*)(ws + L.ck);
        int32_t* cv = (int32_t*)(ws + L.cv);
        uint32_t* ck2 = (uint32_t*)(ws + L.ck2);
        int32_t* cv2 = (int32_t*)(ws + L.cv2);
        const FrontLayout F = front_layout(C, N, tile_w, tile_h);
        if (F.rank_bytes > 0) {  // ranked placement: the front call left the prefix matrix and the cell offsets in its workspace
            FG_REQUIRE(front_workspace && (size_t)front_workspace_bytes >= F.total,
                       "fg_render_back needs the workspace fg_render_front ran with");
            const unsigned char* fws = (const unsigned char*)front_workspace;
            if ((e = fg_bin_ranked_emit(C, N, order, means2d, radii, tile_size, tile_w, tile_h, fws + F.rank,
                                        (int64_t)F.rank_bytes, (const int32_t*)(fws + F.coff), cv, stream)))
                return e;
            if ((e = fg_bin_fine(C, N, n_coarse, (const int32_t*)(fws + F.coff), cv, means2d, radii, tile_size, tile_w,
                                 tile_h, isect_offsets, flatten_ids, stream)))
                return e;
        } else {
            if ((e = fg_bin_coarse_emit(C, N, order, means2d, radii, coarse_off, tile_size, tile_w, tile_h, ck, cv, stream)))
                return e;
            int bits = 1;
            while ((1ll << bits) < (long long)C * cw * chh) ++bits;
            int sel = 0;
            if ((e = fg_radix_sort_pairs_u32_u32(n_coarse, ck, (uint32_t*)cv, ck2, (uint32_t*)cv2, bits, ws + L.sort,
                                                 (int64_t)(L.total - L.sort), &sel, stream)))
                return e;
            const uint32_t* ks = sel ? ck2 : ck;
            const int32_t* vs = sel ? cv2 : cv;
            int32_t* coff = (int32_t*)(ws + L.coff);
            if ((e = fg_isect_offsets_tiles(n_coarse, ks, C, cw, chh, coff, stream))) return e;
            if ((e = fg_bin_fine(C, N, n_coarse, coff, vs, means2d, radii, tile_size, tile_w, tile_h, isect_offsets,
                                 flatten_ids, stream)))
                return e;
        }
    }
    if (CH == 0) return FG_OK;  // lists only: the caller composites later with fg_rasterize_fwd
    return fg_rasterize_fwd(C, N, CH, width, height, tile_size, means2d, conics, feat, opacities, backgrounds,
                            flow_affine, flow_ch0, split, ed_channel, opac_shared, isect_offsets, flatten_ids, n_isects,
                            render, render2, alphas, last_ids, stream);
}
