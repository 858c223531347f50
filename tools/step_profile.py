#!/usr/bin/env python
"""Where does one bench step spend host and device time?  (torch.profiler, GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from freegaussian_b200.rendering import rasterization
from freegaussian_b200.dist import DensificationStats

args = bench.parse_args()
dev = torch.device("cuda", 0)
n, W, H = bench.WORKLOADS[args.workload]
sc = bench.build_scene(args.workload, args.recipe, dev, 8)
d = sc.to(dev)
params = [d.means, d.quats, d.scales, d.opacities, d.sh, d.means_next]
for p in params: p.requires_grad_(True)
wr, wf = (t.to(dev) for t in bench.loss_weights(H, W, 1))
stats = DensificationStats(n, dev)

def step(i, with_stats=True):
    for p in params: p.grad = None
    vm, K = d.viewmats[i % 8: i % 8 + 1], d.Ks[i % 8: i % 8 + 1]
    render, alpha, meta = rasterization(d.means, d.quats, d.scales, d.opacities, d.sh, vm, K, W, H, packed=False,
                                        render_mode="RGB+ED", sh_degree=3, absgrad=True, means_next=d.means_next)
    meta["means2d"].retain_grad()
    loss = (render * wr).sum() + (meta["flow"] * wf).sum()
    loss.backward()
    if with_stats:
        stats.accumulate_local(meta["radii"], meta["means2d"].absgrad, H, W)

for i in range(5): step(i)
torch.cuda.synchronize()
for label, ws in (("full step", True), ("without stats", False)):
    t0 = time.perf_counter()
    for i in range(20): step(i, ws)
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"{label}: host issue time {t_cpu/20*1e3:.3f} ms/step, wall {t_all/20*1e3:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=60))

# chronological device timeline of the last profiled step: kernel, duration, idle gap before it
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
starts = [i for i, e in enumerate(evs) if "project_fwd" in e.name]
if starts:
    last = evs[starts[-1]:]
    t_prev = last[0].time_range.start
    busy = 0.0
    print("\n# device timeline of one step (us): gap-before  duration  kernel")
    for e in last:
        gap = e.time_range.start - t_prev
        dur = e.time_range.end - e.time_range.start
        busy += dur
        print(f"{gap:9.1f} {dur:9.1f}  {e.name[:90]}")
        t_prev = max(t_prev, e.time_range.end)
    span = t_prev - last[0].time_range.start
    print(f"# span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us")
