#!/usr/bin/env python
"""Generate the committed golden fixtures.  Run from the repo root:  python tests/golden/make_golden.py

* knn_*.npz    -- inputs + outputs of the REFERENCE's own k-NN call
                  (sklearn NearestNeighbors, freegaussian_model.py:293-311), via oracle.knn.reference_knn.
* render_*.npz -- seeded scene + outputs of the oracle restatement of the render path.  The reference
                  ships no golden vectors for this path and gsplat cannot be installed here (SURVEY.md
                  8(c)), so these pin the ORACLE against regressions; parity with gsplat stays unpinned.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import knn as oknn  # noqa: E402
from oracle import render as O  # noqa: E402
from util import small_scene  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def knn_fixture(name, x, k):
    d, i = oknn.reference_knn(x, k)
    np.savez_compressed(os.path.join(OUT, f"knn_{name}.npz"), x=x, k=k, dist=d, idx=i.astype(np.int32))


def render_fixture(name, n, w, h, views, seed, **kw):
    sc = small_scene(n, w, h, views=views, seed=seed)
    r, a, m = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, sc.Ks, w, h,
                              means_next=sc.means_next, **kw)
    np.savez_compressed(
        os.path.join(OUT, f"render_{name}.npz"),
        means=sc.means.numpy(), quats=sc.quats.numpy(), scales=sc.scales.numpy(), opacities=sc.opacities.numpy(),
        sh=sc.sh.numpy(), viewmats=sc.viewmats.numpy(), Ks=sc.Ks.numpy(), means_next=sc.means_next.numpy(),
        width=w, height=h, kwargs=np.array(repr(kw)),
        render=r.numpy().astype(np.float32), alpha=a.numpy().astype(np.float32), flow=m["flow"].numpy(),
        radii=m["radii"].numpy(), means2d=m["means2d"].numpy(), depths=m["depths"].numpy(),
        conics=m["conics"].numpy(), flatten_ids=m["flatten_ids"].numpy(), isect_ids=m["isect_ids"].numpy(),
        isect_offsets=m["isect_offsets"].numpy(), tiles_per_gauss=m["tiles_per_gauss"].numpy())


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    knn_fixture("uniform_k3", ((rng.random((4000, 3), dtype=np.float32) - 0.5) * 6.0), 3)  # model.py:155,158 recipe
    knn_fixture("uniform_k16", ((rng.random((3000, 3), dtype=np.float32) - 0.5) * 6.0), 16)
    pl = (rng.random((3000, 3), dtype=np.float32) - 0.5) * 6.0
    pl[:1500, 2] = 1.5
    knn_fixture("planes_k3", pl, 3)
    render_fixture("rgbed_sh3", 600, 64, 48, 2, 21, sh_degree=3, render_mode="RGB+ED")
    render_fixture("rgb_sh0_aa", 500, 56, 40, 1, 22, sh_degree=0, render_mode="RGB", rasterize_mode="antialiased")
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
