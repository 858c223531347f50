#!/usr/bin/env python
"""fg_knn_f32 at the BASELINE cfg5 shape (k = 16 over 3 M points), twice: the ncu target of tools/capture_profiles.sh."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freegaussian_b200.knn import k_nearest  # noqa: E402

g = torch.Generator().manual_seed(11)
x = ((torch.rand(3_000_000, 3, generator=g) - 0.5) * 6.0).cuda()
for _ in range(2):
    d, i = k_nearest(x, 16)
torch.cuda.synchronize()
print(float(d.mean()))
