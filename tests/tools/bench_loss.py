#!/usr/bin/env python
"""Fused blend+L1+SSIM loss vs the torch formulation the reference uses (pytorch_msssim-style conv2d ops), 1080p, GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from freegaussian_b200.losses import blend_l1_ssim_loss
from oracle import loss as OL  # used here only as "the torch formulation" to time beside the kernel

H, W = 1080, 1920
g = torch.Generator().manual_seed(0)
alpha = torch.rand(1, H, W, 1, generator=g).cuda()
render = (torch.rand(1, H, W, 4, generator=g).cuda() * alpha).requires_grad_(True)
alpha.requires_grad_(True)
bg = torch.tensor([0.149, 0.1647, 0.2157]).cuda()
gt = torch.rand(H, W, 3, generator=g).cuda()


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fused():
    render.grad = None; alpha.grad = None
    blend_l1_ssim_loss(render, alpha, bg, gt).backward()


def torch_ops():
    render.grad = None; alpha.grad = None
    OL.blend_l1_ssim_loss(render[0], alpha[0], bg, gt).backward()


a, b = timeit(fused), timeit(torch_ops)
lf = float(blend_l1_ssim_loss(render, alpha, bg, gt)); lt = float(OL.blend_l1_ssim_loss(render[0], alpha[0], bg, gt))
print(f"1080p blend+L1+SSIM fwd+bwd: fused {a:.3f} ms, torch ops {b:.3f} ms ({b/a:.1f}x); loss {lf:.6f} vs {lt:.6f}")
