#!/usr/bin/env python
"""Kernel-level breakdown of one deformation-network step (torch profiler, CUDA activities):  python tools/prof_deform.py [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freegaussian_b200.deform import DeformNetwork  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
torch.manual_seed(1)
net = DeformNetwork(is_blender=True).cuda()  # nn.Linear's default initialisation
g = torch.Generator().manual_seed(0)
m = ((torch.rand(n, 3, generator=g) - 0.5) * 6).cuda().requires_grad_(True)
s = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.005).cuda().requires_grad_(True)
q = torch.randn(n, 4, generator=g).cuda().requires_grad_(True)
t = torch.tensor([[0.3]]).cuda().expand(n, -1)


def step():
    a, b, c = net.deform_gaussians(m, s, q, t)
    (a.sum() + b.sum() + c.sum()).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
