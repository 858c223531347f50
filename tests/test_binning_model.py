"""CPU models of the arithmetic the ranked tile-binning kernels rely on (csrc/binning.cu): the algorithm and its integer
identities restated in numpy and checked against the definition (a stable sort by cell).  The kernels themselves are checked
bit for bit against the oracle's lists on the GPU (tests/test_gpu_stages.py); these tests pin the REASONING."""
import numpy as np

CK = 4
CHUNK = 512


def _rects(rng, n, cw, chh, big_every=17):
    x0 = rng.integers(0, cw, n)
    y0 = rng.integers(0, chh, n)
    w = rng.integers(1, 4, n)
    h = rng.integers(1, 4, n)
    w[::big_every] = rng.integers(1, cw + 1, len(w[::big_every]))  # some splats cover most of the image
    h[::big_every] = rng.integers(1, chh + 1, len(h[::big_every]))
    x1 = np.minimum(x0 + w, cw)
    y1 = np.minimum(y0 + h, chh)
    return x0, x1, y0, y1


def test_ranked_placement_equals_the_stable_sort_by_cell():
    """position = cell offset + pairs of earlier chunks in the cell + earlier splats of the chunk touching the cell."""
    rng = np.random.default_rng(0)
    cw, chh, n = 30, 17, 5000
    x0, x1, y0, y1 = _rects(rng, n, cw, chh)
    n_cells = cw * chh
    # definition: emit (cell, splat) pairs in splat (= depth) order, stable sort by cell
    keys, vals = [], []
    for i in range(n):
        for y in range(y0[i], y1[i]):
            for x in range(x0[i], x1[i]):
                keys.append(y * cw + x)
                vals.append(i)
    keys, vals = np.array(keys), np.array(vals)
    want = vals[np.argsort(keys, kind="stable")]
    # bin_count_cells: per (chunk, cell) counts through a difference grid in cell space
    n_chunks = -(-n // CHUNK)
    mat = np.zeros((n_chunks, n_cells), np.int64)
    for c in range(n_chunks):
        g = np.zeros((chh + 1, cw + 1), np.int64)
        for i in range(c * CHUNK, min(n, (c + 1) * CHUNK)):
            g[y0[i], x0[i]] += 1
            g[y0[i], x1[i]] -= 1
            g[y1[i], x0[i]] -= 1
            g[y1[i], x1[i]] += 1
        mat[c] = g.cumsum(1).cumsum(0)[:chh, :cw].reshape(-1)
    assert mat.sum() == len(keys)
    # cell_scan: exclusive prefix over chunks per cell, exclusive prefix of the totals over cells
    totals = mat.sum(0)
    cell_offsets = np.concatenate([[0], np.cumsum(totals)])
    prefix = np.cumsum(mat, 0) - mat
    # ranked_emit: rank inside the chunk = earlier splats of the chunk that touch the same cell
    got = np.full(len(keys), -1, np.int64)
    for c in range(n_chunks):
        seen = np.zeros(n_cells, np.int64)
        for i in range(c * CHUNK, min(n, (c + 1) * CHUNK)):
            for y in range(y0[i], y1[i]):
                for x in range(x0[i], x1[i]):
                    cell = y * cw + x
                    got[cell_offsets[cell] + prefix[c, cell] + seen[cell]] = i
                    seen[cell] += 1
    assert np.array_equal(got, want)


def test_division_by_multiply_shift_is_exact_on_the_kernels_domain():
    """WarpRect::cell: i / w == (i * ceil(2^20 / w)) >> 20 for every rectangle of at most 1024 cells (i < n <= 1024, w <= n)."""
    for w in range(1, 1025):
        inv = ((1 << 20) + w - 1) // w
        i = np.arange(0, 1024, dtype=np.uint64)
        assert np.array_equal((i * np.uint64(inv)) >> np.uint64(20), i // np.uint64(w)), w
        assert int(i.max()) * inv < 2**32  # the product fits the kernel's 32-bit multiply


def test_tile_mask_by_multiplication():
    """cell_tile_mask / fine_bin_kernel: the 4-bit column pattern replicated into rows [y0, y1) of a 4x4 cell."""
    for x0 in range(4):
        for x1 in range(x0 + 1, 5):
            for y0 in range(4):
                for y1 in range(y0 + 1, 5):
                    cols = ((1 << x1) - 1) & ~((1 << x0) - 1)
                    rows = ((1 << (y1 * CK)) - 1) & ~((1 << (y0 * CK)) - 1)
                    want = 0
                    for y in range(y0, y1):
                        want |= cols << (y * CK)
                    assert (cols * 0x1111) & rows == want


def test_reciprocal_scaling_is_exact_for_power_of_two_tiles():
    """tile_rect: x * (1 / 16) == x / 16 in float32 for every finite x (scaling by a power of two is exact)."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * np.float32(3000.0),
                        rng.integers(-2**31, 2**31, 200000).astype(np.int32).view(np.float32)])
    x = x[np.isfinite(x)]
    inv = np.float32(1.0) / np.float32(16.0)
    assert np.array_equal((x * inv).view(np.uint32), (x / np.float32(16.0)).view(np.uint32))


def test_compact_row_of_a_published_colour_gradient():
    """csrc/exchange.cu (sh_bwd_views_kernel) finds the compact row of the visible pair (view c, Gaussian n) as
    prefix[c][n >> 5] + popcount(mask[c][n >> 5] & ((1 << (n & 31)) - 1)), where prefix[c][w] is the exclusive scan of the
    visibility at the first bit of word w (project.cu publishes exactly that).  Model: it equals the exclusive scan itself."""
    rng = np.random.default_rng(2)
    C, N = 3, 1000
    vis = rng.random((C, N)) < 0.3
    flat = vis.reshape(-1)
    excl = np.cumsum(flat) - flat                     # fg_pack_plan: exclusive scan over c*N+n
    words = (N + 31) // 32
    mask = np.zeros((C, words), np.uint64)
    prefix = np.zeros((C, words), np.int64)
    for c in range(C):
        for n in range(N):
            if vis[c, n]:
                mask[c, n >> 5] |= np.uint64(1) << np.uint64(n & 31)
        for w in range(words):
            prefix[c, w] = excl[c * N + 32 * w]       # what lane 0 of each warp stores
    for c in range(C):
        for n in range(N):
            if vis[c, n]:
                below = int(mask[c, n >> 5]) & ((1 << (n & 31)) - 1)
                assert prefix[c, n >> 5] + bin(below).count("1") == excl[c * N + n]
