#!/usr/bin/env python
"""The exchange kernels on ONE GPU (one-rank process group + FG_XCHG_SOLO=1): cfg3 scene, 8 views on this rank, three
backward passes through ViewShardedExchange -- the ncu target of tools/capture_profiles.sh for sh_bwd_views_kernel and
allreduce_kernel (their NVLink behaviour is measured by bench.py at N > 1)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FG_XCHG_SOLO"] = "1"
import bench  # noqa: E402
from freegaussian_b200.dist import ViewShardedExchange  # noqa: E402
from freegaussian_b200.rendering import rasterization  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", world_size=1, rank=0, device_id=dev)
V = int(os.environ.get("FG_SOLO_VIEWS", "8"))
sc = bench.build_scene("cfg3", "trained_like", dev, V).to(dev)
n, W, H = bench.WORKLOADS["cfg3"]
W, H = W // 2, H // 2  # 8 views of 960x540: the same per-Gaussian work as 8 ranks x 1 view, a quarter of the pixels
Ks = sc.Ks.clone()
Ks[:, :2] *= 0.5
xc = ViewShardedExchange().install()
params = [sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.means_next]
for p in params:
    p.requires_grad_(True)
for _ in range(3):
    for p in params:
        p.grad = None
    r, a, m = rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats, Ks, W, H, packed=False,
                            render_mode="RGB+ED", sh_degree=3, absgrad=True, means_next=sc.means_next)
    (r.sum() + m["flow"].sum()).backward()
torch.cuda.synchronize()
print("ok", float(sc.sh.grad.abs().sum()))
dist.destroy_process_group()
