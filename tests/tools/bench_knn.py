#!/usr/bin/env python
"""k-NN kernel timing (cfg5 shape: k=16 over 3 M points) next to the reference's sklearn call on a sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from freegaussian_b200.knn import k_nearest
from oracle import knn as OK

for n, k in [(1_000_000, 3), (3_000_000, 3), (3_000_000, 16)]:
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(n, 3, generator=g) - 0.5) * 6.0
    xd = x.cuda()
    k_nearest(xd, k); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); d, i = k_nearest(xd, k); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ns = 200_000
    t0 = time.perf_counter(); rd, ri = OK.reference_knn(x[:ns].numpy(), k); t_ref = time.perf_counter() - t0
    d2, i2 = k_nearest(xd[:ns].contiguous(), k)
    same = np.array_equal(d2.cpu().numpy(), rd)
    print(f"n={n} k={k}: GPU {min(ts)*1e3:.1f} ms ({n/min(ts)/1e6:.1f} M queries/s); sklearn on {ns} pts: {t_ref:.2f} s "
          f"({ns/t_ref/1e6:.3f} M queries/s, 1 core); bit-exact distances on the sample: {same}")
