"""k-NN oracle (CPU): the plain float64 restatement the CUDA kernel follows is pinned to the reference's own
sklearn call (freegaussian_model.py:293-311) and to fixtures generated from it."""
from pathlib import Path

import numpy as np
import pytest

from oracle import knn as OK

GOLD = Path(__file__).parent / "golden"


@pytest.mark.parametrize("name", ["uniform_k3", "uniform_k16", "planes_k3"])
def test_brute_matches_reference_fixture(name):
    z = np.load(GOLD / f"knn_{name}.npz")
    d, i = OK.brute_knn(z["x"], int(z["k"]))
    assert np.array_equal(d, z["dist"]), "distances must be bit-identical to sklearn's"
    assert np.array_equal(i, z["idx"].astype(np.int64))


def test_reference_call_reproduces_fixture():
    z = np.load(GOLD / "knn_uniform_k3.npz")
    d, i = OK.reference_knn(z["x"], int(z["k"]))
    assert np.array_equal(d, z["dist"]) and np.array_equal(i, z["idx"].astype(np.int64))


def test_self_is_dropped_and_sorted():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(500, 3)).astype(np.float32)
    d, i = OK.brute_knn(x, 5)
    assert (i != np.arange(500)[:, None]).all()
    assert (np.diff(d, axis=1) >= 0).all() and (d > 0).all()
