"""Host-side policies of freegaussian_b200.rendering that need no GPU."""


def test_list_capacity_is_stable_under_drift():
    """rendering._list_capacity: a need that creeps upward changes the buffer size only once per ~20 % of growth."""
    from freegaussian_b200.rendering import _LIST_QUANTUM, _list_capacity

    cap, changes, need = 0, 0, 40_000_000
    for _ in range(2000):
        new = _list_capacity(cap, need)
        assert new >= need and new % _LIST_QUANTUM == 0
        changes += new != cap
        cap = new
        need += 4000  # +20 % over the run
    assert changes <= 2
    assert _list_capacity(cap, 1000) == cap  # never shrinks here (the 64-small-calls rule in _Project does that)


def _lib():
    from freegaussian_b200 import _build, _lib
    _build.build()
    return _lib.lib()


def test_published_block_layout_is_consistent():
    """fg_xchg_pub_layout / fg_xchg_pub_bytes (host arithmetic of csrc/exchange.cu): parts in order, 16-byte aligned (the
    pull kernel moves 16-byte units), non-overlapping, large enough for the worst case (every (view, Gaussian) visible)."""
    import ctypes as C
    L = _lib()
    for V, N in [(1, 1), (1, 1_000_000), (4, 3_000_000), (8, 12345), (3, 31), (2, 32), (5, 33)]:
        o = [C.c_int64(-1) for _ in range(4)]
        words = C.c_int32(-1)
        assert L.fg_xchg_pub_layout(V, N, *[C.byref(x) for x in o], C.byref(words)) == 0
        nnz_off, mask_off, prefix_off, rgb_off = (int(x.value) for x in o)
        total = int(L.fg_xchg_pub_bytes(V, N))
        assert words.value >= (N + 31) // 32
        assert 16 * V <= nnz_off and nnz_off + 4 <= mask_off          # camera centres [V,4] f32, then the count
        assert mask_off + 4 * V * words.value <= prefix_off
        assert prefix_off + 4 * V * words.value <= rgb_off
        assert rgb_off + 12 * V * N <= total
        assert all(x % 16 == 0 for x in (mask_off, prefix_off, rgb_off, total))


def test_ranked_binning_applies_up_to_1024_coarse_cells():
    """fg_bin_ranked_workspace_bytes: non-zero exactly when C * ceil(tile_w/4) * ceil(tile_h/4) <= 1024, and large enough for
    the (chunk of 512 slots) x cell matrix."""
    import ctypes as C
    L = _lib()
    for Cn, W, H in [(1, 1920, 1080), (2, 1920, 1080), (3, 1920, 1080), (1, 2704, 2028), (1, 960, 540), (8, 128, 128),
                     (1, 16, 16), (4, 333, 190)]:
        tw, th = -(-W // 16), -(-H // 16)
        cw, ch = C.c_int(0), C.c_int(0)
        assert L.fg_bin_coarse_dims(tw, th, C.byref(cw), C.byref(ch)) == 0
        assert (cw.value, ch.value) == (-(-tw // 4), -(-th // 4))
        cells = Cn * cw.value * ch.value
        for N in (0, 1, 511, 512, 513, 1_000_000):
            got = int(L.fg_bin_ranked_workspace_bytes(Cn, N, tw, th))
            if cells > 1024:
                assert got == 0
            else:
                chunks = -(-(Cn * N) // 512)
                assert got >= 4 * chunks * cells + 4 * cells + 4
        # the front workspace carries it, plus the cell offsets
        assert int(L.fg_render_front_workspace_bytes(Cn, 1000, tw, th)) >= int(L.fg_bin_ranked_workspace_bytes(Cn, 1000, tw, th))


def test_ranked_entry_points_refuse_what_they_cannot_do():
    """The ranked binning calls validate before they touch the device: too many coarse cells, missing or short workspaces
    come back as error codes with a message (no GPU needed to see them)."""
    L = _lib()
    one = 1  # any non-NULL address: the checks below fail before a pointer is followed
    assert L.fg_bin_count_cells(1, 10, one, one, one, 16, 200, 200, one, one, 1 << 30, None) != 0
    assert b"too many coarse cells" in L.fg_last_error()
    need = int(L.fg_bin_ranked_workspace_bytes(1, 10, 120, 68))
    assert need > 0
    assert L.fg_bin_count_cells(1, 10, one, one, one, 16, 120, 68, one, one, need - 1, None) != 0
    assert b"workspace too small" in L.fg_last_error()
    assert L.fg_bin_cell_scan(1, 10, 120, 68, None, one, need, one, one, None) != 0
    assert L.fg_bin_ranked_emit(1, 10, one, one, one, 16, 120, 68, None, need, one, one, None) != 0
