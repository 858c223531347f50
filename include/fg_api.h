/* freegaussian_b200 -- C ABI of the B200-native splat-render hot path.
 *
 * This is the drop-in boundary below the Python `rasterization(...)` call that
 * FreeGaussian makes at freegaussian/freegaussian_model.py:847-868 and
 * freegaussian/freegaussian_control_model.py:158-179.  The
 * reference binds that call to the un-vendored gsplat package; each entry point below
 * names the gsplat stage and the reference line it stands behind.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and no
 *     entry point synchronises the host unless it says so.
 *   - every function returns 0 on success or an FG_ERR_* code; the message of the last
 *     failure on the calling thread is returned by fg_last_error().  Nothing throws.
 *   - arrays are dense row-major float32 / int32 / int64 with the shapes given.
 *   - C = cameras, N = Gaussians, CH = composited channels, M = tile intersections.
 */
#ifndef FG_API_H_
#define FG_API_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_OK 0
#define FG_ERR_INVALID 1   /* bad argument (shape, range, null) */
#define FG_ERR_CUDA 2      /* a CUDA runtime call or launch failed */
#define FG_ERR_WORKSPACE 3 /* workspace too small */

#define FG_MAX_CHANNELS 8 /* channels composited in one pass; wider inputs are chunked by the host */

const char* fg_last_error(void);
/* ABI version; bumped on any signature change. */
int fg_abi_version(void);
/* Number of kernels launched by this library in this process so far (bench.py `gpu_launches`). */
long long fg_launch_count(void);

/* Run-time switches for A/B measurements: "fwd_two_pixels" = 0 (default: the one-pixel forward kernel) | 1 (two pixels
 * per thread, packed FP32; measured slower, kept as the documented experiment; both produce the same images);
 * "xchg_ar_blocks" (default 64) / "xchg_pull_blocks" (default 4): CTAs of the all-reduce kernel / per peer of the
 * pull kernel of the multi-GPU exchange.  FG_ERR_INVALID for an unknown name. */
int fg_set_option(const char* name, int value);

/* Measured FP32 FMA peak of this GPU in TFLOP/s (a short FFMA microbenchmark; synchronises).
 * Denominator for the FP32-pipe-bound compositing kernels' roofline in bench.py. */
int fg_measure_fp32_tflops(double* tflops_host, void* stream);

/* ---- (1) fused projection + EWA covariance + culling + SH->RGB (+ flow features) -------
 * Replaces gsplat `fully_fused_projection` + `spherical_harmonics` + the depth/flow
 * channel concatenation inside `rasterization` as called at freegaussian_model.py:847-868
 * (arguments: near_plane=0.01, far_plane=1e10, eps2d=0.3, radius_clip=0, :859-867).
 *
 * Inputs   means[N,3] quats[N,4] (w,x,y,z, any norm) scales[N,3] (linear)
 *          viewmats[C,4,4] rigid world->camera (utils.py:162-179)  Ks[C,3,3]
 *          sh_coeffs[N,sh_bases,3] with sh_degree in 0..3 evaluated, or sh_degree=-1 (none)
 *          means_next[N,3] (frame t+1; NULL = no flow).  flow_cov!=0 additionally uses
 *          quats_next/scales_next (NULL = reuse frame t) and writes flow_affine.
 * Outputs  radii[C,N] int32 (0 = culled)  means2d[C,N,2]  depths[C,N]  conics[C,N,3]
 *          compensations[C,N] (NULL unless rasterize_mode="antialiased")
 *          feat[C,N,feat_stride]: rgb at channel rgb_off (if sh_degree>=0), depth at
 *          depth_off (if >=0), flow (mu2d(t+1)-mu2d(t)) at flow_off (if >=0).
 *          flow_affine[C,N,4] = B(t+1) B(t)^-1 - I, row-major (covariance flow mode).
 *          tiles_per_gauss[C,N] int32 = number of 16x16 tiles each splat touches.
 */
int fg_project_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                   const float* viewmats, const float* Ks, int width, int height, float eps2d,
                   float near_plane, float far_plane, float radius_clip, int tile_size,
                   int sh_degree, int sh_bases, const float* sh_coeffs, const float* means_next,
                   const float* quats_next, const float* scales_next, int flow_cov, int32_t* radii,
                   float* means2d, float* depths, float* conics, float* compensations, float* feat,
                   int feat_stride, int rgb_off, int depth_off, int flow_off, float* flow_affine,
                   int32_t* tiles_per_gauss, void* stream);

/* Optional output of fg_project_bwd for the view-sharded multi-GPU exchange (section (5) below): instead of writing
 * v_sh rows (pass v_sh = NULL), the SH kernel PUBLISHES, for its own C views: the camera centres campos[C,4] followed by
 * one uint32 nnz; a visibility bit mask mask[C,words] (bit n of view c = radii[c,n] > 0; words >= ceil(N/32)); per mask
 * word the compact row of its first visible splat, prefix[C,words]; and the colour gradient rgb[nnz,3] of the visible
 * (view, Gaussian) pairs with the max(rgb+0.5,0) clamp mask applied, compact, in ascending c*N+n order.  `offsets` /
 * `nnz` are the exclusive scan of (radii > 0) over c*N+n and its total (fg_pack_plan).  The four output pointers
 * normally point into the rank's block of symmetric memory (fg_xchg_pub_layout gives the byte offsets). */
typedef struct {
    float* campos;
    uint32_t* mask;
    uint32_t* prefix;
    float* rgb;
    const int32_t* offsets;
    const int64_t* nnz;
    int32_t words;
    int32_t phase; /* 0: the whole backward; 1: only the SH kernel (publishes, writes the direction term into v_means);
                      2: only the geometry kernel (adds to the v_means a phase-1 call left) -- lets the caller start the
                      cross-rank SH summation on another stream while the geometry kernel runs */
} fg_project_bwd_pub;

/* VJP of fg_project_fwd (gsplat `fully_fused_projection_bwd` + `spherical_harmonics` bwd).
 * Any of v_means2d / v_depths / v_conics / v_compensations / v_feat / v_flow_affine may be
 * NULL (= zero).  Outputs are fully overwritten (no pre-zeroing needed):
 *   v_means[N,3] v_quats[N,4] v_scales[N,3] v_sh[N,sh_bases,3] (if sh_degree>=0)
 *   v_means_next[N,3] (if means_next) v_quats_next[N,4] v_scales_next[N,3] (if flow_cov and given)
 * Gradients of the C cameras are summed inside one thread per Gaussian (deterministic).
 * `feat` is the forward output (or NULL): with it and 16-byte aligned rows of a multiple of four
 * bases, the SH half runs as its own streaming kernel (clamp mask read from feat's rgb channels).
 */
int fg_project_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                   const float* viewmats, const float* Ks, int width, int height, float eps2d,
                   float near_plane, float far_plane, float radius_clip, int sh_degree,
                   int sh_bases, const float* sh_coeffs, const float* means_next,
                   const float* quats_next, const float* scales_next, int flow_cov,
                   const int32_t* radii, const float* v_means2d, const float* v_depths,
                   const float* v_conics, const float* v_compensations, const float* v_feat,
                   const float* feat, int feat_stride, int rgb_off, int depth_off, int flow_off,
                   const float* v_flow_affine, float* v_means, float* v_quats, float* v_scales,
                   float* v_sh, float* v_means_next, float* v_quats_next, float* v_scales_next,
                   const fg_project_bwd_pub* pub, void* stream);

/* ---- (2) tile intersection: scan, emission, sort, offsets ------------------------------
 * Replaces gsplat `isect_tiles` (+cumsum), `cub::DeviceRadixSort::SortPairs`, and
 * `isect_offset_encode` (tile_size=16 from freegaussian_model.py:806).
 */

/* Exclusive prefix sum of counts[n] -> offsets[n] (int32), total -> *total (int64, device).
 * workspace: fg_scan_workspace_bytes(n) bytes. */
int64_t fg_scan_workspace_bytes(int64_t n);
int fg_exclusive_scan_i32(int64_t n, const int32_t* counts, int32_t* offsets, int64_t* total,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* Emit (key64, flatten_id) for every (splat, tile) pair in flattened (c*N+n) order:
 *   key = c << (32+tile_bits) | (ty*tile_w+tx) << 32 | float_bits(depth),
 *   tile_bits = floor(log2(tile_w*tile_h)) + 1, value = c*N+n.
 * `offsets` is the exclusive scan of tiles_per_gauss. */
int fg_isect_emit(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                  const int32_t* offsets, int tile_size, int tile_w, int tile_h, int64_t* isect_ids,
                  int32_t* flatten_ids, void* stream);

/* Stable LSD radix sort of (uint64 key, uint32 value) pairs on key bits [0, end_bit).
 * Hand-written onesweep (chained-scan, decoupled look-back); the result is the unique
 * stable order, bit-identical to cub::DeviceRadixSort::SortPairs.  The two buffer pairs are
 * used as ping-pong storage (both are clobbered); *result_in_out (host int) is set to 1 if
 * the sorted data ended in keys_out/vals_out, 0 if it ended in keys_in/vals_in. */
int64_t fg_radix_sort_workspace_bytes(int64_t n);
int fg_radix_sort_pairs_u64_u32(int64_t n, uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out,
                                uint32_t* vals_out, int end_bit, void* workspace,
                                int64_t workspace_bytes, int* result_in_out, void* stream);
/* Same for 32-bit keys (depth-only and tile-only sorts of the two-level path). */
int fg_radix_sort_pairs_u32_u32(int64_t n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                                uint32_t* vals_out, int end_bit, void* workspace,
                                int64_t workspace_bytes, int* result_in_out, void* stream);

/* offsets[c,ty,tx] (int32, C*tile_h*tile_w entries) = lower bound of that tile in the sorted keys. */
int fg_isect_offsets(int64_t n_isects, const int64_t* sorted_isect_ids, int C, int tile_w, int tile_h,
                     int32_t* offsets, void* stream);

/* Two-level variant of the same sort (identical resulting order, ~4x less sort traffic):
 * (a) stable-sort the (c,n) splats once by depth bits (fg_isect_depth_keys + the u32 radix
 * sort; splats touching no tile get key 0xffffffff), (b) emit their tiles in that order with
 * 32-bit keys cam*tile_w*tile_h + tile (fg_gather_i32 of the counts, scan, fg_isect_emit_tiles),
 * (c) stable-sort by tile key (2 radix passes instead of 6).  Ties at equal (camera, tile,
 * depth) keep ascending c*N+n, exactly like the stable 64-bit sort. */
int fg_isect_depth_keys(int64_t total, const float* depths, const int32_t* tiles_per_gauss,
                        uint32_t* keys, uint32_t* vals, void* stream);
/* The same order for the VISIBLE splats only, in one call (the default `binned` path): order[0 .. n_visible) = flat ids
 * (c*N+n) of the splats with tiles_per_gauss > 0, stably sorted by depth bits; order[n_visible .. total) = -1 (the
 * binning stages stop there); *n_visible_dev (int64, device) = their number.  Keys and the four digit histograms come
 * from one kernel, the first radix pass drops the culled splats while it sorts, the other three run over the visible
 * ones only (their count is read on the device).  workspace: fg_depth_sort_workspace_bytes(total). */
int64_t fg_depth_sort_workspace_bytes(int64_t total);
int fg_depth_sort_visible(int64_t total, const float* depths, const int32_t* tiles_per_gauss, int32_t* order,
                          int64_t* n_visible_dev, void* workspace, int64_t workspace_bytes, void* stream);
int fg_gather_i32(int64_t n, const int32_t* src, const int32_t* idx, int32_t* dst, void* stream);
int fg_isect_emit_tiles(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                        const int32_t* offsets, int tile_size, int tile_w, int tile_h, uint32_t* tile_keys,
                        int32_t* flatten_ids, void* stream);
int fg_isect_offsets_tiles(int64_t n_isects, const uint32_t* sorted_tile_keys, int C, int tile_w, int tile_h,
                           int32_t* offsets, void* stream);
/* Rebuild the reference's sorted 64-bit keys (gsplat meta["isect_ids"]) from the two-level result. */
int fg_isect_ids_from_tiles(int64_t n_isects, const uint32_t* sorted_tile_keys, const int32_t* flatten_ids,
                            const float* depths, int tile_w, int tile_h, int64_t* isect_ids, void* stream);

/* Hierarchical binning: the same per-tile lists and offsets with no sort over the M tile
 * intersections at all (default path of rendering.py; csrc/binning.cu explains the steps).
 *   order[C*N]       flat ids (c*N+n) stably sorted by depth (fg_isect_depth_keys + u32 sort)
 *   fg_bin_count     corner increments of every splat's tile rectangle into diff_grid
 *                    [C,tile_h+1,tile_w+1] (zeroed inside) + coarse-cell count per splat (in `order`)
 *   fg_bin_tile_scan diff_grid -> per-tile counts (in place) -> isect_offsets, *total = M (device)
 *   fg_bin_coarse_emit  (cell key, flat id) pairs in depth order; coarse_off = exclusive scan of the counts;
 *                    cells are 4x4 tiles, key = cam*cw*ch + cy*cw + cx (fg_bin_coarse_dims gives cw, ch)
 *   fg_bin_fine      after a stable sort of those pairs by key and fg_isect_offsets_tiles-style
 *                    offsets per cell: append every splat to the lists of the tiles it overlaps */
int fg_bin_coarse_dims(int tile_w, int tile_h, int* cw, int* ch);
int fg_bin_count(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                 int tile_size, int tile_w, int tile_h, int32_t* diff_grid, int32_t* coarse_cnt, void* stream);
int64_t fg_bin_tile_scan_workspace_bytes(int C, int tile_w, int tile_h);
int fg_bin_tile_scan(int C, int tile_w, int tile_h, int32_t* diff_grid, int32_t* isect_offsets,
                     int64_t* total, void* workspace, int64_t workspace_bytes, void* stream);
int fg_bin_coarse_emit(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii,
                       const int32_t* coarse_off, int tile_size, int tile_w, int tile_h,
                       uint32_t* coarse_keys, int32_t* coarse_vals, void* stream);
/* Ranked placement of the (splat, cell) pairs: the same cell-grouped, depth-ordered `coarse_vals` as fg_bin_coarse_emit
 * followed by a stable sort by cell, computed without a sort.  Applies when C * cw * ch <= 1024 coarse cells (one or two
 * 1080p views); fg_bin_ranked_workspace_bytes returns 0 otherwise and the emit + sort path is the one to use.
 *   fg_bin_count_cells  fg_bin_count's corner increments + per (chunk of 512 depth-ordered slots, cell) pair counts
 *   fg_bin_cell_scan    per cell: exclusive prefix over the chunks; cell_offsets[n_cells + 1] (= fg_bin_fine's
 *                       coarse_offsets, then the total), *n_coarse = Mc (device).  n_visible: device count from
 *                       fg_depth_sort_visible
 *   fg_bin_ranked_emit  coarse_vals[Mc]: position = cell offset + chunk prefix + rank inside the chunk (bitmaps) */
int64_t fg_bin_ranked_workspace_bytes(int C, int N, int tile_w, int tile_h);
int fg_bin_count_cells(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii, int tile_size,
                       int tile_w, int tile_h, int32_t* diff_grid, void* ranked_workspace, int64_t ranked_workspace_bytes,
                       void* stream);
int fg_bin_cell_scan(int C, int N, int tile_w, int tile_h, const int64_t* n_visible, void* ranked_workspace,
                     int64_t ranked_workspace_bytes, int32_t* cell_offsets, int64_t* n_coarse, void* stream);
int fg_bin_ranked_emit(int C, int N, const int32_t* order, const float* means2d, const int32_t* radii, int tile_size,
                       int tile_w, int tile_h, const void* ranked_workspace, int64_t ranked_workspace_bytes,
                       const int32_t* cell_offsets, int32_t* coarse_vals, void* stream);
int fg_bin_fine(int C, int N, int64_t n_coarse, const int32_t* coarse_offsets,
                const int32_t* coarse_vals_sorted, const float* means2d, const int32_t* radii, int tile_size,
                int tile_w, int tile_h, const int32_t* isect_offsets, int32_t* flatten_ids, void* stream);

/* ---- (3) per-tile front-to-back alpha compositing, forward and backward ----------------
 * Replaces gsplat `rasterize_to_pixels` fwd/bwd and the wrapper's "ED" normalisation.  One
 * pass composites all CH channels (RGB + depth + flow).  alpha = min(0.999, opacity*exp(-sigma));
 * skip alpha < 1/255; stop when T(1-alpha) <= 1e-4.
 *   feat[C*N,CH]; opacities[C*N], or [N] shared by all cameras when opac_shared != 0
 *   backgrounds[C,CH] or NULL (blended in the epilogue: out += T * bg)
 *   flow_affine[C*N,4] or NULL: channels flow_ch0, flow_ch0+1 get + A_g (p - mu_g) per pixel
 *   split: channels [0,split) are written to render[C,H,W,split], channels [split,CH) to
 *          render2[C,H,W,CH-split] (split == CH: render2 unused) -- RGB(D) and flow come back as
 *          two dense images without a slicing pass
 *   ed_channel (< split, or -1): that channel is divided by max(alpha,1e-10) (render_mode "ED")
 *   alphas[C,H,W]  last_ids[C,H,W] int32 (index into the sorted list of the last composited splat)
 */
int fg_rasterize_fwd(int C, int N, int CH, int width, int height, int tile_size, const float* means2d,
                     const float* conics, const float* feat, const float* opacities,
                     const float* backgrounds, const float* flow_affine, int flow_ch0, int split,
                     int ed_channel, int opac_shared, const int32_t* isect_offsets,
                     const int32_t* flatten_ids, int64_t n_isects, float* render, float* render2,
                     float* alphas, int32_t* last_ids, void* stream);

/* Backward.  v_* outputs must be zero-initialised by the caller (they are accumulated with
 * atomics); v_opacities is [N] when opac_shared.  v_render / v_render2 / v_alphas may be NULL
 * (= zero).  `render` (the forward output) is only read when ed_channel >= 0.  v_means2d_abs may
 * be NULL (absgrad=False); v_flow_affine NULL unless flow_affine. */
int fg_rasterize_bwd(int C, int N, int CH, int width, int height, int tile_size, const float* means2d,
                     const float* conics, const float* feat, const float* opacities,
                     const float* backgrounds, const float* flow_affine, int flow_ch0, int split,
                     int ed_channel, int opac_shared, const int32_t* isect_offsets,
                     const int32_t* flatten_ids, int64_t n_isects, const float* render, const float* alphas,
                     const int32_t* last_ids, const float* v_render, const float* v_render2,
                     const float* v_alphas, float* v_means2d, float* v_means2d_abs, float* v_conics,
                     float* v_feat, float* v_opacities, float* v_flow_affine, void* stream);

/* Densification statistics of freegaussian_model.py:369-392 folded over this rank's C views in
 * one pass: grad_norm[n] += sum_c vis ? |absgrad[c,n]|_2 : 0; vis_count[n] += sum_c vis;
 * max_size[n] = max(max_size[n], max_c radii[c,n] / max(H,W)). */
int fg_densify_stats(int C, int N, const int32_t* radii, const float* absgrad, float inv_max_hw,
                     float* grad_norm, float* vis_count, float* max_size, void* stream);

/* ---- (3b) the forward pass in two calls ------------------------------------------------------
 * fg_render_front = fg_project_fwd + depth sort + fg_bin_count + fg_bin_tile_scan + the coarse scan,
 * then the ONE host synchronisation of the pass: counts_host[0] = M (tile intersections),
 * counts_host[1] = Mc (coarse pairs), counts_host[2] = 1 if the lists were built already: when the caller
 * passes a flatten_ids buffer of a guessed capacity >= M (and a back workspace large enough for Mc), the
 * list-building half of fg_render_back is enqueued right here, without a round trip through the host mirror
 * (pass NULL / 0 to opt out).  Otherwise the caller allocates flatten_ids[M] and calls
 * fg_render_back = coarse pairs grouped by cell (ranked placement when it applies, else emit + sort + cell offsets) +
 * fg_bin_fine + fg_rasterize_fwd.  `front_workspace` is the workspace fg_render_front ran with, untouched since (the ranked
 * placement reads its prefix matrix and cell offsets from it).
 * Arguments are those of the granular entry points; `order`, `coarse_off` are [C*N] int32.
 * fg_render_back with CH == 0 builds the tile lists only (flatten_ids) and skips compositing: the host
 * mirror calls it right after the sync so the GPU is busy again while Python assembles the compositing call.
 */
int64_t fg_render_front_workspace_bytes(int C, int N, int tile_w, int tile_h);
int fg_render_front(int C, int N, const float* means, const float* quats, const float* scales,
                    const float* viewmats, const float* Ks, int width, int height, float eps2d,
                    float near_plane, float far_plane, float radius_clip, int tile_size, int sh_degree,
                    int sh_bases, const float* sh_coeffs, const float* means_next, const float* quats_next,
                    const float* scales_next, int flow_cov, int32_t* radii, float* means2d, float* depths,
                    float* conics, float* compensations, float* feat, int feat_stride, int rgb_off,
                    int depth_off, int flow_off, float* flow_affine, int32_t* tiles_per_gauss, int32_t* order,
                    int32_t* isect_offsets, int32_t* coarse_off, int64_t* counts_host, void* workspace,
                    int64_t workspace_bytes, int32_t* flatten_ids, int64_t flatten_capacity,
                    void* back_workspace, int64_t back_workspace_bytes, void* stream);
int64_t fg_render_back_workspace_bytes(int C, int tile_w, int tile_h, int64_t n_coarse);
int fg_render_back(int C, int N, int64_t n_isects, int64_t n_coarse, const int32_t* order,
                   const int32_t* coarse_off, const float* means2d, const int32_t* radii, int tile_size,
                   const int32_t* isect_offsets, int32_t* flatten_ids, void* workspace, int64_t workspace_bytes,
                   const void* front_workspace, int64_t front_workspace_bytes,
                   int CH, int width, int height, const float* conics, const float* feat, const float* opacities,
                   const float* backgrounds, const float* flow_affine, int flow_ch0, int split, int ed_channel,
                   int opac_shared, float* render, float* render2, float* alphas, int32_t* last_ids, void* stream);

/* ---- (4) exact k-nearest neighbours ----------------------------------------------------
 * Replaces FreeGaussianModel.k_nearest_sklearn (freegaussian_model.py:293-311):
 * sklearn NearestNeighbors(n_neighbors=k+1, metric="euclidean") of the set against itself
 * with the self column dropped.  Distances are float64 sqrt(dx^2+dy^2+dz^2) in x,y,z order
 * (bit-identical to sklearn's kd-tree), returned as float32; indices int32; ascending
 * (distance, index) order.  workspace: fg_knn_workspace_bytes(n). */
int64_t fg_knn_workspace_bytes(int64_t n);
int fg_knn_f32(int64_t n, const float* points /*[n,3]*/, int k, float* out_dist /*[n,k]*/,
               int32_t* out_idx /*[n,k]*/, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- (5) preprocess: attribute-mask assignment ---------------------------------------------
 * Loop body of preprocess/knn_gaussian.py:116-132 after the packed "ED" render (:93-113): a visible
 * Gaussian whose truncated projected centre is inside the image and whose depth is consistent with
 * the rendered depth D there (-0.1 D < D - z < D) gets gaussian_masks[gaussian_id, a] = 1 for every
 * attribute a with atrb_masks[y,x,a] && mask_valids[a].  Masks are bytes (torch.bool). */
int fg_assign_masks(int64_t nnz, const float* means2d, const float* depths, const int64_t* gaussian_ids,
                    const float* depth_img, int width, int height, const uint8_t* atrb_masks,
                    const uint8_t* mask_valids, int n_attr, uint8_t* gaussian_masks, void* stream);

/* ---- (6) fused blend + clamp + L1 + SSIM loss ------------------------------------------------
 * pred = clamp(render[..., :3] + (1 - alpha) bg, 0, 1) (freegaussian_model.py:876-877);
 * loss = (1-l) mean|gt - pred| + l (1 - SSIM(gt, pred)) (freegaussian_model.py:965-981; SSIM of
 * pytorch_msssim: 11-tap Gaussian, sigma 1.5, valid separable convolution).
 * fwd: sums[0] = (1-l) mean|gt-pred|, sums[1] = l * mean SSIM (doubles, device); `partial` =
 * fg_l1_ssim_workspace_floats(W,H) floats kept for the backward.  bwd: v_render[H,W,render_stride]
 * (channels >= 3 zeroed), v_alpha[H,W], scaled by the device scalar *v_loss.
 * mask[H,W] (float, or NULL): gt and pred are both multiplied by it before the loss (freegaussian_model.py:957-963). */
int64_t fg_l1_ssim_workspace_floats(int width, int height);
int fg_l1_ssim_fwd(int width, int height, int render_stride, const float* render, const float* alpha,
                   const float* background, const float* gt, const float* mask, float ssim_lambda, float* partial,
                   double* sums, void* stream);
int fg_l1_ssim_bwd(int width, int height, int render_stride, const float* render, const float* alpha,
                   const float* background, const float* gt, const float* mask, float ssim_lambda, const float* partial,
                   const float* v_loss, float* v_render, float* v_alpha, void* stream);

/* Depth fix-up of freegaussian_model.py:884-886:
 *   depth[i] = alpha[i] > 0 ? render[i, channel] : max over ALL pixels of render[:, channel]
 * (the expected-depth channel of an "RGB+ED" render; the maximum is a constant for the backward, as the reference
 * detaches it).  max_ws: one uint32 of scratch.  bwd writes the full v_render[n_pixels, render_stride]
 * (zero outside `channel` and where alpha == 0). */
int fg_depth_fixup_fwd(int64_t n_pixels, const float* render, int render_stride, int channel, const float* alpha,
                       float* depth, uint32_t* max_ws, void* stream);
int fg_depth_fixup_bwd(int64_t n_pixels, const float* alpha, const float* v_depth, int render_stride, int channel,
                       float* v_render, void* stream);

/* ---- (7) "next" row: optimizer step and refinement -----------------------
 * fg_adam_step: one launch steps every Gaussian parameter group the reference gives its own
 * torch.optim.Adam (freegaussian_config.py:48-75: eps=1e-15, betas (0.9, 0.999), no weight decay),
 * following torch's single-tensor update operation by operation.  A segment whose rows hold two
 * groups (the [N,16,3] SH tensor = features_dc ++ features_rest, freegaussian_model.py:801) takes a
 * learning rate per column range.  `first` is the index of param[0] inside the whole tensor, so a
 * rank can step only its shard after a reduce-scatter of the gradient arena.  `step` counts from 1.
 * betas and eps are doubles because torch derives 1-beta and the bias corrections in double.
 */
#define FG_ADAM_MAX_SEGMENTS 8
typedef struct fg_adam_segment {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t n;       /* elements of this (shard of the) tensor */
    int64_t first;   /* index of param[0] in the whole tensor (0 unless sharded) */
    int32_t row_len; /* floats per Gaussian row; 0 = a single learning rate */
    int32_t split;   /* columns [0,split) use lr, [split,row_len) use lr_rest */
    float lr;
    float lr_rest;
} fg_adam_segment;
int fg_adam_step(int n_segments, const fg_adam_segment* segments_host, int step, double beta1, double beta2,
                 double eps, void* stream);

/* Refinement (freegaussian_model.py:404-491 with split_gaussians :513-556, dup_gaussians :558-567,
 * cull_gaussians :493-511 and the Adam-state surgery :313-367) as a plan / map / gather pipeline.
 * scales are log-space [N,3], opacities logit-space [N] as the model stores them.
 *
 * fg_refine_plan: per Gaussian the reference's masks
 *     high   = (grad_norm/vis_count) * 0.5 * max(W,H) > densify_grad_thresh            (:421-422)
 *     split  = (max exp(scale) > densify_size_thresh & high) | (use_screen & max_size > split_screen_size)
 *     dup    = (max exp(scale') <= densify_size_thresh) & high, scale' = the scale after split_gaussians
 *              rescaled its parents in place (:536 runs before :430-431)
 *     cull   = sigmoid(opacity) < cull_alpha_thresh | split | (cull_big & (max exp(scale) > cull_scale_thresh
 *              | (use_screen & max_size > cull_screen_size)))                          (:499-511)
 *   (children of a split carry scale - log 1.6 and max_size 0, duplicates the parent's scale and
 *   max_size 0, and pass through the same cull test), then an exclusive scan.  With densify == 0
 *   only the cull runs (:467-468) and grad_norm / vis_count / max_size may be NULL.
 *   plan[4N+1] int32 receives the scanned positions (last = total); counts[4] (device) = kept originals, parents
 *   whose children survive, surviving duplicates, split parents.
 * fg_refine_map: src[n_out] = parent row of every output row in the reference's order
 *   (kept originals, then children sample-major, then duplicates); sample_row[n_out] = row of the
 *   torch.randn((n_split_samples * n_split, 3)) draw (:519) a child consumes, -1 for copies, -2 for the
 *   duplicate of a parent that was also split (it copies the rescaled scale).
 * fg_refine_gather: out[d,:] = in[src[d],:] for every listed array; arrays with zero_new != 0
 *   (Adam moments, :344-357) get zeros in rows d >= n_keep.
 * fg_refine_children: for rows with sample_row >= 0: means = parent mean + R(q/|q|) (exp(scale) * z)
 *   (:520-526), scales = log(exp(scale) / 1.6) (:535); rows with sample_row == -2: scales only.
 */
typedef struct fg_refine_config {
    float densify_grad_thresh, densify_size_thresh, split_screen_size;
    float cull_alpha_thresh, cull_scale_thresh, cull_screen_size;
    float max_dim;           /* max(W, H) of the last render (:421) */
    int32_t n_split_samples; /* :427 */
    int32_t use_screen;      /* step < stop_screen_size_at */
    int32_t cull_big;        /* step > refine_every * reset_alpha_every (:505) */
    int32_t densify;         /* 0 = cull only */
} fg_refine_config;
typedef struct fg_refine_array {
    const float* in;
    float* out;
    int32_t row_floats;
    int32_t zero_new;
} fg_refine_array;
#define FG_REFINE_MAX_ARRAYS 24
int64_t fg_refine_workspace_bytes(int64_t N);
int fg_refine_plan(int64_t N, const float* scales, const float* opacities, const float* grad_norm,
                   const float* vis_count, const float* max_size, const fg_refine_config* config_host,
                   int32_t* plan, int64_t* counts, void* workspace, int64_t workspace_bytes, void* stream);
int fg_refine_map(int64_t N, const int32_t* plan, const int64_t* counts, int n_split_samples, int64_t n_out,
                  int32_t* src, int32_t* sample_row, void* stream);
int fg_refine_gather(int64_t n_out, int64_t n_keep, const int32_t* src, int n_arrays,
                     const fg_refine_array* arrays_host, void* stream);
int fg_refine_children(int64_t n_out, int64_t n_keep, int64_t n_children, const int32_t* src,
                       const int32_t* sample_row, const float* samples, const float* means,
                       const float* quats, const float* scales, float* means_out, float* scales_out,
                       void* stream);

/* ---- deformation network ----------------------------------------------------------
 * The MLP FreeGaussian evaluates on every Gaussian right before each render call once warm-up is over:
 * FreeGaussianDeformableModel.forward (freegaussian_model.py:1091-1114; 8 x Linear(256)+ReLU, skip after
 * layer 4, four linear heads) and the application of its outputs (freegaussian_model.py:836-845).
 * These entry points replace the torch.nn.Linear / F.relu / torch.cat / exp_se3 / torch.bmm calls of those lines.
 *
 * Numerics of the tensor-core kernel: error-compensated 3xTF32.  Each fp32 operand x is used as hi + lo, hi = x rounded
 * to the nearest tf32 value and lo = x - hi; every k-step accumulates A_lo.W_hi + A_hi.W_lo + A_hi.W_hi in fp32, which is
 * fp32-accurate (the reference computes these layers in fp32).  Activations are plain fp32 row-major arrays whose row
 * length is a multiple of 32 floats (zero padded); they are split on chip.  Weights are passed pre-split (fg_mlp_pack).
 *
 * fg_mlp_linear: out[M, n_out] = epilogue( [A0 | A1] . W^T ),  A0 [M,k0], A1 [M,k1] (k1 may be 0), W [n_out, k0+k1] as
 *   w_hi / w_lo.
 *   FG_MLP_RELU    n_out = 256: out = max(. + bias, 0); mask_out[M, 8] uint32, bit j of word c = (column 32c+j > 0)
 *                  (nn.Linear + F.relu, :1097-1099)
 *   FG_MLP_LINEAR  n_out = FG_MLP_HEAD_LD or 128: out = . + bias   (the heads, :1103-1112 / :1144; 128: the gradient
 *                  of the embedding, A0 / A1 = dL/d(output) of the two layers that read it, zero bias)
 *   FG_MLP_DGRAD   n_out = 256: out = mask_in bit ? . : 0   (data gradient of Linear + ReLU: A0 = dL/d(output of
 *                  layer l), W = W_l^T, mask_in = the mask_out of layer l-1)
 * fg_mlp_pack: copies column ranges of reference-layout weights ([rows, src_ld] row-major) into the padded operand
 *   buffers (optionally transposed), splitting hi / lo (dst_lo may be NULL).  One launch for the whole table.
 * fg_deform_embed: E[n,:] = [embed(x[n]) | embed(x2[n]) if x2 | t_emb[t_ch] | 0], [N, ld], ld a multiple of 32, with
 *   embed(p) = [p, sin(p 2^0), cos(p 2^0), ..., sin(p 2^(multires-1)), cos(..)]  (Embedder.embed, utils.py:27-56; the
 *   torch.cat at freegaussian_model.py:1096 (x, time) and :1137 (x, value; stage 2)).
 * fg_deform_embed_bwd: dx[N,3] = VJP of embed(x) for de[N, ld] (columns of x2 / t_emb ignored).
 * fg_deform_apply_fwd: head[N, FG_MLP_HEAD_LD] = (branch_w 3 | branch_v 3 | gaussian_rotation 4 | gaussian_scaling 3 | 0)
 *   -> theta = |w|, screw axis (w, v) / theta + 1e-5, exp_se3 (utils.py:137-159), means' = R means + p,
 *   scales' = exp(scales_log) + d_scaling, quats' = quats / |quats| + d_rotation   (freegaussian_model.py:841-845).
 * fg_deform_apply_bwd: the VJP of the above: v_head [N, FG_MLP_HEAD_LD], v_means (the direct path; the network saw
 *   means.detach()), v_scales_log, v_quats.
 */
#define FG_MLP_EMBED_LD 96
#define FG_MLP_HEAD_LD 32
#define FG_MLP_PACK_MAX_SEGMENTS 32
#define FG_MLP_RELU 0
#define FG_MLP_LINEAR 1
#define FG_MLP_DGRAD 2
typedef struct fg_mlp_pack_segment {
    const float* src;
    float* dst_hi;
    float* dst_lo;
    int32_t src_ld, src_col0, rows, cols; /* source block [rows, cols] at column src_col0 */
    int32_t dst_ld, dst_col0;             /* destination row length and first column */
    int32_t transpose;                    /* != 0: dst[c, r] = src[r, c] */
} fg_mlp_pack_segment;
int fg_mlp_linear(int mode, int64_t M, int n_out, const float* a0, int k0, const float* a1, int k1, const float* w_hi,
                  const float* w_lo, const float* bias, const uint32_t* mask_in, float* out, uint32_t* mask_out, void* stream);
/* fg_mlp_wgrad: dw[256, ld_dw] columns col0 .. col0 + k_in - 1 += dz^T . a, and (db != NULL) db[256] += column sums of dz;
 * dz [N, 256], a [N, k_in] with k_in = 256, 128, FG_MLP_EMBED_LD or FG_MLP_HEAD_LD, fp32 row-major.  The weight / bias gradient
 * of nn.Linear (dL/dW = dz^T a, dL/db = sum_n dz); with dz = the last hidden layer and a = the head gradient it yields
 * the TRANSPOSED head weight gradient.  3xTF32 on chip, split-K over row ranges, results ADDED with
 * red.global (zero dw / db first; the order of the additions is not deterministic). */
int fg_mlp_wgrad(int64_t N, const float* dz, const float* a, int k_in, float* dw, int ld_dw, int col0, float* db, void* stream);
int fg_mlp_pack(int n_segments, const fg_mlp_pack_segment* segments_host, void* stream);
int fg_deform_embed(int64_t N, const float* x, const float* x2, const float* t_emb, int t_ch, int multires, int ld, float* e,
                    void* stream);
int fg_deform_embed_bwd(int64_t N, const float* x, const float* de, int multires, int ld, float* dx, void* stream);
int fg_deform_apply_fwd(int64_t N, const float* head, const float* means, const float* scales_log, const float* quats,
                        float* means_out, float* scales_out, float* quats_out, void* stream);
int fg_deform_apply_bwd(int64_t N, const float* head, const float* means, const float* scales_log, const float* quats,
                        const float* v_means_out, const float* v_scales_out, const float* v_quats_out, float* v_head,
                        float* v_means, float* v_scales_log, float* v_quats, void* stream);


/* packed=True layout of the render call (preprocess/knn_gaussian.py:93-113 and the other preprocess
 * scripts): compact every per-(camera, Gaussian) tensor to the visible pairs in ascending c*N+n order.
 * fg_pack_plan: offsets[C*N] = exclusive scan of (radii > 0), *nnz_dev = number of visible pairs (int64, device).
 * fg_pack_gather: row offsets[i] of every *_p output = row i of the input for each visible i = c*N+n; camera_ids /
 * gaussian_ids (int64) = c / n.  fg_pack_remap: flatten_ids[m] = offsets[flatten_ids[m]] in place. */
int64_t fg_pack_workspace_bytes(int64_t total);
int fg_pack_plan(int64_t total, const int32_t* radii, int32_t* offsets, int64_t* nnz_dev, void* workspace,
                 int64_t workspace_bytes, void* stream);
int fg_pack_gather(int C, int N, int CH, const int32_t* radii, const int32_t* offsets, const float* means2d,
                   const float* depths, const float* conics, const float* feat, const float* opacities, int opac_shared,
                   const float* flow_affine, int32_t* radii_p, float* means2d_p, float* depths_p, float* conics_p,
                   float* feat_p, float* opac_p, float* flow_affine_p, int64_t* camera_ids, int64_t* gaussian_ids,
                   void* stream);
int fg_pack_remap(int64_t M, const int32_t* offsets, int32_t* flatten_ids, void* stream);

/* Host-sync-free sparse backward of the networks (csrc/rows.cu).  fg_rows_active: idx[0 .. count) = indices (ascending) of
 * the rows of g[N, ld] with a non-zero element, idx[count .. N) = 0, *count_dev = count (int64, device).  fg_rows_gather:
 * dst[r] = r < *count_dev ? src[idx[r]] : 0 for r < M, rows of row_bytes (a multiple of 16): M is a capacity the host chose
 * without knowing the count; zero rows contribute nothing to the products of the backward. */
int64_t fg_rows_workspace_bytes(int64_t N);
int fg_rows_active(int64_t N, const float* g, int ld, int32_t* idx, int64_t* count_dev, void* workspace,
                   int64_t workspace_bytes, void* stream);
int fg_rows_gather(int64_t M, const int32_t* idx, const int64_t* count_dev, const void* src, int row_bytes, void* dst,
                   void* stream);
/* Time branch of FreeGaussianDeformableModel (freegaussian_model.py:1066-1071, 1094-1096) for the ONE time value of a call:
 * emb[in_ch] = [t, sin(2^k t), cos(2^k t)]_k (utils.py:27-56); with w1 != NULL: h[hidden] = relu(W1 emb + b1),
 * out[out_ch] = W2 h + b2 (`timenet`).  bwd: gradients of the four timenet tensors from g_out[out_ch]. */
int fg_time_branch_fwd(const float* t, int multires, int in_ch, int hidden, int out_ch, const float* w1, const float* b1,
                       const float* w2, const float* b2, float* emb, float* h, float* out, void* stream);
int fg_time_branch_bwd(int in_ch, int hidden, int out_ch, const float* emb, const float* h, const float* w2,
                       const float* g_out, float* dw1, float* db1, float* dw2, float* db2, void* stream);

/* ---- (5) view-sharded multi-GPU exchange over NVLink peer memory / NVSwitch multicast ------------------
 * BASELINE.json north_star item 5; the reference itself is single-GPU, so these have no reference
 * counterpart: they make `loss.backward()` on every rank return the gradient one process rendering all views would.
 * One process per GPU; every rank owns a block of SYMMETRIC memory (same size on every rank, peer-mapped into every
 * rank's address space, optionally also mapped through an NVSwitch multicast object).  `buf[r]` / `flags[r]` are rank
 * r's block and flag area as seen from THIS rank; `mc` is the multicast alias of `buf` (NULL: no NVLS, the kernels use
 * peer loads / stores instead).  The flag area is FG_XCHG_FLAG_BYTES bytes, zeroed once before first use.  `epoch`
 * must grow by one per call (all ranks pass the same value); calls must be issued in the same order on every rank. */
#define FG_XCHG_MAX_RANKS 16
#define FG_XCHG_FLAG_BYTES 65536
typedef struct {
    int32_t world, rank;
    void* buf[FG_XCHG_MAX_RANKS];
    void* mc;
    void* flags[FG_XCHG_MAX_RANKS];
} fg_xchg_peers;

/* Capacity in bytes of one rank's published block for V views of N Gaussians (campos + nnz | mask | prefix | rgb, see
 * fg_project_bwd_pub), and the byte offsets of its parts. */
int64_t fg_xchg_pub_bytes(int V, int N);
int fg_xchg_pub_layout(int V, int N, int64_t* nnz_off, int64_t* mask_off, int64_t* prefix_off, int64_t* rgb_off,
                       int32_t* words);
/* Cross-rank barrier on the stream: returns (on the device) once every rank's stream has reached it. */
int fg_xchg_barrier(const fg_xchg_peers* peers_host, uint32_t epoch, void* stream);
/* In-place all-reduce(SUM) of n_floats (multiple of 4) at offset_bytes (multiple of 16) of the symmetric buffer.
 * Two-shot inside ONE kernel: rank r multimem.ld_reduce's the r-th slice (summed by the switch), multimem.st's it to
 * every rank; start_barrier != 0 first waits until every rank has reached the call (its inputs are complete), and the
 * kernel ends with a barrier, so that on return every rank holds the full result.  Same value on every rank. */
int fg_xchg_allreduce_f32(const fg_xchg_peers* peers_host, int64_t offset_bytes, int64_t n_floats, uint32_t epoch,
                          int start_barrier, void* stream);
/* v_sh[N,sh_bases,3] = sum over ALL ranks' views of basis(dir(n, view)) x published colour gradient.  First PULLS every
 * peer's published range (fixed part + 12 nnz bytes, at pub_offset_bytes of its symmetric buffer) over NVLink with coalesced
 * 16-byte loads into `staging` (world blocks of staging_stride >= fg_xchg_pub_bytes(V, N) bytes, ordinary device memory),
 * then rebuilds the rows from local memory; rows of invisible splats never travel.  Fixed summation order (rank, view):
 * bit-identical on every rank.  Call after a barrier that follows the publishing fg_project_bwd on every rank. */
int fg_xchg_sh_bwd_views(const fg_xchg_peers* peers_host, int64_t pub_offset_bytes, int V, int N, int sh_degree,
                         int sh_bases, const float* means, float* v_sh, void* staging, int64_t staging_stride,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FG_API_H_ */
