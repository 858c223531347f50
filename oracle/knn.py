"""k-NN oracle for ``FreeGaussianModel.k_nearest_sklearn``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY PINNED: ``reference_knn`` *is* the reference's implementation -- the same
scikit-learn call with the same arguments as ``freegaussian_model.py:293-311``
(``NearestNeighbors(n_neighbors=k+1, algorithm="auto", metric="euclidean")``,
column 0 dropped) -- and scikit-learn is importable in this image.  ``brute_knn``
is the plain restatement (float64 ``sqrt(dx^2 + dy^2 + dz^2)`` in x,y,z order, stable
ascending (distance, index) order) that the CUDA kernel follows; tests pin it to
``reference_knn`` and to the committed fixtures in ``tests/golden/knn_*.npz``.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np


def reference_knn(x: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """The reference call (``freegaussian_model.py:305-311``), minus its lossy fp32 index cast.

    Returns distances [N,k] float32 and indices [N,k] int64 (self column removed).
    """
    from sklearn.neighbors import NearestNeighbors

    model = NearestNeighbors(n_neighbors=k + 1, algorithm="auto", metric="euclidean").fit(x)
    dist, idx = model.kneighbors(x)
    return dist[:, 1:].astype(np.float32), idx[:, 1:].astype(np.int64)


def brute_knn(x: np.ndarray, k: int, block: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """Exact k-NN by exhaustive float64 distances; ties broken by ascending index.

    Same contract as :func:`reference_knn` on duplicate-free inputs.
    """
    x64 = np.asarray(x, dtype=np.float32).astype(np.float64)
    n = x64.shape[0]
    out_d = np.empty((n, k), np.float32)
    out_i = np.empty((n, k), np.int64)
    for s in range(0, n, block):
        q = x64[s : s + block]
        dx = q[:, None, 0] - x64[None, :, 0]
        dy = q[:, None, 1] - x64[None, :, 1]
        dz = q[:, None, 2] - x64[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        order = np.argsort(d2, axis=1, kind="stable")[:, : k + 1]
        rows = np.arange(q.shape[0])[:, None]
        dist = np.sqrt(d2[rows, order])
        # drop self: it is column 0 for duplicate-free inputs
        out_d[s : s + block] = dist[:, 1:].astype(np.float32)
        out_i[s : s + block] = order[:, 1:]
    return out_d, out_i
