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
