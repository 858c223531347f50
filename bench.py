#!/usr/bin/env python
"""bench.py -- fwd+bwd MPix/s (RGB+depth+flow) of the splat-render hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (config.workload = "cfg3"): 1 M Gaussians, 1920x1080, SH degree 3, render_mode
"RGB+ED" + rendered flow (6 composited channels), one view per GPU per step (view-sharded,
weak scaling), scalar loss = sum(render * w_rgbd) + sum(flow * w_flow), backward to all
Gaussian parameters; at N>1 the step ends with the NCCL all-reduce of the parameter gradients
and the densification statistics (freegaussian_b200/dist.py).  Synthetic "trained-like" scene
(SURVEY.md 8(d)); working set (236 MB of parameters + ~0.5 GB of intersection buffers) is far
larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (n_gaussians, width, height)
    "cfg1": (10_000, 128, 128),
    "cfg2": (300_000, 960, 540),
    "cfg3": (1_000_000, 1920, 1080),
    "cfg4": (3_000_000, 2704, 2028),
}
N_VIEW_POOL = 8  # distinct cameras cycled through per rank
REFINE_EVERY = 100  # config/sim/base.yaml:22 (refine_every)
CPU_CROP = 128   # the CPU arm renders a CPU_CROP x CPU_CROP centre crop of the same frame


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--recipe", default="trained_like", choices=["trained_like", "init_like"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--views-per-gpu", type=int, default=1, help="views each rank renders per step (cfg4: 4)")
    ap.add_argument("--no-train-iter", action="store_true", help="skip the stage-1 training-iteration section (train_iter)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = published colour gradients + in-switch all-reduce inside the backward (csrc/exchange.cu); "
                         "'nccl' = one NCCL all-reduce over the dense 236 B/Gaussian arena (round-1 path, kept for comparison)")
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ workload
def build_scene(workload: str, recipe: str, device, n_views: int):
    from freegaussian_b200.knn import k_nearest
    from freegaussian_b200.scenes import make_scene
    n, w, h = WORKLOADS[workload]
    knn3 = lambda m: k_nearest(m.to(device), 3)[0].cpu()  # the product KNN kernel seeds the scales (model.py:158)
    sc = make_scene(n, w, h, n_views=n_views, recipe=recipe, seed=0, knn3=knn3)
    return sc


def loss_weights(h: int, w: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, h, w, 4, generator=g), torch.rand(1, h, w, 2, generator=g) * 0.1


def run_ours(args):
    import torch.distributed as dist
    from freegaussian_b200 import _lib
    from freegaussian_b200 import rendering
    from freegaussian_b200.dist import DensificationStats, ViewShardedExchange, exchange
    from freegaussian_b200.rendering import rasterization

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL_DEBUG=VERSION prints a banner on stdout; stdout is the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    from freegaussian_b200 import _build
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    L = _lib.lib()
    xchg = None
    if world > 1 and args.exchange == "peer":
        # the exchange runs inside the projection backward over NVLink peer memory (csrc/exchange.cu)
        xchg = ViewShardedExchange().install()

    n, W, H = WORKLOADS[args.workload]
    sc = build_scene(args.workload, args.recipe, dev, N_VIEW_POOL * world)
    d = sc.to(dev)
    params = [d.means, d.quats, d.scales, d.opacities, d.sh, d.means_next]
    for p in params:
        p.requires_grad_(True)
    V = args.views_per_gpu
    my_views = list(range(rank, N_VIEW_POOL * world, world))  # dist.shard_views
    # per-step host inputs (pinned): cameras + the loss weight images standing in for GT rgb/depth/flow
    pick = lambda t, j: torch.cat([t[my_views[(j + k) % len(my_views)]][None] for k in range(V)], 0)
    host_vm = [pick(sc.viewmats, j).clone().pin_memory() for j in range(len(my_views))]
    host_K = [pick(sc.Ks, j).clone().pin_memory() for j in range(len(my_views))]
    w_rgbd_h, w_flow_h = loss_weights(H, W, 1 + rank)
    w_rgbd_h, w_flow_h = w_rgbd_h.repeat(V, 1, 1, 1), w_flow_h.repeat(V, 1, 1, 1)
    w_rgbd_h, w_flow_h = w_rgbd_h.pin_memory(), w_flow_h.pin_memory()
    w_rgbd, w_flow = w_rgbd_h.to(dev), w_flow_h.to(dev)
    dev_vm = [v.to(dev) for v in host_vm]
    dev_K = [k.to(dev) for k in host_K]
    stats = DensificationStats(n, dev)
    info = {}
    state = {"it": 0}

    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}
    loss_ring = torch.full((64,), float("nan")).pin_memory()
    quantiles = {}

    # two device-side staging sets, filled alternately on the copy stream (a dataloader's double buffer)
    stage_bufs = [(torch.empty_like(dev_vm[0]), torch.empty_like(dev_K[0]), torch.empty_like(w_rgbd),
                   torch.empty_like(w_flow)) for _ in range(2)]
    consumed = [None, None]  # event: the step that read staging set k has finished with it

    def stage_inputs(i: int):
        """H2D copy of step i's host inputs (pinned) on the copy stream: the usual input prefetch --
        step i+1's camera and target images travel while step i computes, all inside the timed region."""
        j = i % len(my_views)
        k = i % 2
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            bufs = stage_bufs[k]
            for dst, src in zip(bufs, (host_vm[j], host_K[j], w_rgbd_h, w_flow_h)):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (bufs, ev)

    def step(i: int, e2e: bool):
        j = i % len(my_views)
        if e2e:
            if i not in staged:
                stage_inputs(i)
            (vm, K, wr, wf), ev = staged.pop(i)
            torch.cuda.current_stream().wait_event(ev)
            stage_inputs(i + 1)  # prefetch the next step's inputs behind this step's kernels
        else:
            vm, K, wr, wf = dev_vm[j], dev_K[j], w_rgbd, w_flow
        for p in params:
            p.grad = None
        render, alpha, meta = rasterization(d.means, d.quats, d.scales, d.opacities, d.sh, vm, K, W, H,
                                            packed=False, near_plane=0.01, far_plane=1e10, render_mode="RGB+ED",
                                            sh_degree=3, sparse_grad=False, absgrad=True, rasterize_mode="classic",
                                            means_next=d.means_next)
        meta["means2d"].retain_grad()
        loss = (render * wr).sum() + (meta["flow"] * wf).sum()
        loss.backward()
        stats.accumulate_local(meta["radii"], meta["means2d"].absgrad, H, W)
        # exchange step: ONE all-reduce over the flat gradient arena (no-op at N=1).  The densification
        # statistics are accumulated per rank by one kernel per step and reduced across ranks when they
        # are consumed (refine_every = 100 steps in the reference configs) -- same numbers.
        if xchg is None and world > 1:
            with rendering._stage("exchange"):
                exchange([p.grad for p in params])
        state["it"] += 1
        if state["it"] % REFINE_EVERY == 0:
            stats.sync()
        info["meta"] = meta
        if e2e:
            consumed[i % 2] = torch.cuda.Event()
            consumed[i % 2].record()
            # D2H read of the step's result: 4 bytes into a pinned ring, read by the host once the copy has
            # landed (checked at the next step's list-size sync and at the end of the timed region), the way a
            # training loop logs its loss without stalling the launch queue
            loss_ring[i % len(loss_ring)].copy_(loss.detach(), non_blocking=True)
        return None

    def timed(k: int, e2e: bool):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        loss_ring.fill_(float("nan"))
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        e0, e1 = marks[0], marks[-1]
        e0.record()
        for i in range(k):
            step(i, e2e)
            marks[i + 1].record()
        torch.cuda.synchronize()
        if e2e:
            assert bool(torch.isfinite(loss_ring[:min(k, len(loss_ring))]).all()), "a step's loss never reached the host"
        per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(k))
        pct = lambda q: per_step[min(k - 1, int(q * k))]
        quantiles["e2e" if e2e else "dev"] = {"p10": pct(0.10), "p50": pct(0.50), "p90": pct(0.90)}
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # nvidia-smi takes a few hundred ms to deliver its first sample, longer than a timed region: it is started before
    # the warm-up and stopped after the second timed region, so its samples cover warm-up + both timed regions (all under load)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i, False)
    l0 = _lib.launch_count()
    ms_dev = timed(args.steps, False)
    launches = _lib.launch_count() - l0
    step(0, True)
    staged.clear()
    ms_e2e = timed(args.steps, True)
    staged.clear()
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing + roofline (every rank runs the steps -- they contain collectives --
    #      rank 0 records the CUDA-event time of each C-ABI stage)
    out = None
    rendering.stage_timer.enabled = rank == 0
    rendering.stage_timer.reset()
    for i in range(min(args.steps, 8)):
        step(i, False)
    torch.cuda.synchronize()
    if rank == 0:
        stage_ms = {k: statistics.mean(v) for k, v in rendering.stage_timer.summary().items()}
        rendering.stage_timer.enabled = False
        meta = info["meta"]
        M = int(meta["flatten_ids"].numel())
        n_vis = int((meta["radii"] > 0).sum())
        # evaluated (pixel, Gaussian) pairs a pixel must visit: from its tile's list start to its last contributor
        offs = meta["isect_offsets"]
        start_px = offs.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W]
        contributed = meta["last_ids"] >= start_px
        pairs = int(((meta["last_ids"] - start_px + 1).clamp(min=0) * contributed).sum())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        import ctypes
        tf = ctypes.c_double(0)
        L.fg_measure_fp32_tflops(ctypes.byref(tf), torch.cuda.current_stream().cuda_stream)
        fp32_peak = tf.value
        n_tiles = math.ceil(W / 16) * math.ceil(H / 16)
        if rendering.SORT_MODE == "key64":
            cam_bits, tile_bits = 1, int(math.floor(math.log2(n_tiles))) + 1
            passes = (32 + tile_bits + cam_bits + 7) // 8
            sort_bytes, emit_bytes = M * (8 + passes * 24), n * 20 + M * 12
        else:  # two-level: 32-bit tile keys, ceil(log2(tiles)/8) passes; depth sort of the n splats separately
            passes = (max(1, math.ceil(math.log2(n_tiles))) + 7) // 8
            sort_bytes, emit_bytes = M * (4 + passes * 16), n * 24 + M * 8
        algo = {  # algorithmic bytes / flops per launch (DESIGN.md "Kernels")
            "project_fwd": ("hbm", n_vis * 276 + (n - n_vis) * 44),
            "project_bwd": ("hbm", n_vis * 548 + (n - n_vis) * (44 + 4 + 236)),
            "sort": ("hbm", sort_bytes),
            "depth_sort": ("hbm", n * (8 + 8) + n * (4 + 4 * 16)),
            "emit": ("hbm", emit_bytes),
            "bin_count": ("hbm", n * 4 + n_vis * 16 + n * 12),
            "fine_bin": ("hbm", M * 4 + n_vis * 16 * 14),
            "rasterize_fwd": ("fp32", pairs * 24),
            "rasterize_bwd": ("fp32", pairs * 70),
        }
        kernels = {}
        for name, (bound, work) in algo.items():
            if name not in stage_ms:
                continue
            t = stage_ms[name] * 1e-3
            if bound == "hbm":
                ach = work / t / 1e9
                kernels[name] = {"bound": "hbm", "ms": stage_ms[name], "achieved": ach, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": ach / hbm_peak}
            else:
                ach = work / t / 1e12
                kernels[name] = {"bound": "fp32", "ms": stage_ms[name], "achieved": ach, "peak": fp32_peak,
                                 "unit": "TFLOP/s", "frac": ach / fp32_peak if fp32_peak else None}
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
        # captures (profiles/r1_final_*.txt); only valid for the configuration they were taken on
        ncu_traffic = {}
        if args.workload == "cfg3" and args.recipe == "trained_like":
            # project_bwd = sh_bwd_kernel (269.8 MB) + project_bwd_kernel<-1> (135.4 MB), the two launches of the stage
            ncu_traffic = {"rasterize_bwd": 127.7e6, "rasterize_fwd": 24.4e6, "project_bwd": 405.2e6,
                           "project_fwd": 147.7e6, "fine_bin": 100.3e6}
        for name, k in kernels.items():
            k["traffic"] = ncu_traffic.get(name)
        dominant = max(stage_ms, key=stage_ms.get)
        roof = dict(kernels.get(dominant, {}))
        roof.update({"kernel": dominant, "traffic": ncu_traffic.get(dominant),
                     "peak_source": "fg_measure_fp32_tflops (FFMA microbenchmark, this run)" if roof.get("bound") == "fp32" else hbm_src})
        hbm_kernels = {k: v for k, v in kernels.items() if v["bound"] == "hbm"}
        dom_hbm = max(hbm_kernels, key=lambda k: hbm_kernels[k]["ms"]) if hbm_kernels else None

        pix = world * V * W * H
        out = {
            "metric": "fwd+bwd MPix/s (RGB+depth+flow) at 1M Gaussians",
            "value": pix * args.steps / (ms_dev * 1e-3) / 1e6,
            "unit": "MPix/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps,
            "ms_per_step_quantiles": quantiles.get("dev"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n} Gaussians, {W}x{H}, 1 view/GPU/step, SH3, RGB+ED+flow (6 ch), "
                                   f"{args.recipe} scene", "views_per_step": world * V, "l2": "inputs larger than L2 (no flush)",
                       "n_isects": M, "visible": n_vis, "sort_mode": rendering.SORT_MODE, "pairs_per_pixel": pairs / (V * W * H),
                       "parallelism": f"view-sharded dp{world}" if world > 1 else "single GPU"},
            "e2e": {"value": pix * args.steps / (ms_e2e * 1e-3) / 1e6, "unit": "MPix/s",
                    "h2d_bytes_per_step": int(host_vm[0].numel() * 4 + host_K[0].numel() * 4 + w_rgbd_h.numel() * 4 + w_flow_h.numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "ms_per_step_quantiles": quantiles.get("e2e")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "roofline_hbm": dict(hbm_kernels[dom_hbm], kernel=dom_hbm, peak_source=hbm_src) if dom_hbm else None,
            "kernels": kernels,
            "stage_ms": stage_ms,
        }
        if world > 1:
            geo_b = 4 * (n * 14 + 3 * n)  # means 3 + quats 4 + scales 3 + opacity 1 + means_next 3 floats, reduced in the switch
            if xchg is not None:
                out["exchange"] = {"mode": "peer: published colour gradients (12 B per visible (view, Gaussian)) read over NVLink by "
                                           "fg_xchg_sh_bwd_views + in-switch two-shot all-reduce of the geometry gradients",
                                   "multicast": bool(xchg.multicast), "ms": stage_ms.get("exchange"),
                                   "allreduce_bytes": geo_b, "published_bytes_per_rank": n_vis * 12 + n // 8,
                                   "dense_arena_bytes_replaced": 4 * n * 62}
            else:
                out["exchange"] = {"mode": "nccl: one all-reduce over the dense gradient arena", "ms": stage_ms.get("exchange"),
                                   "allreduce_bytes": 4 * n * 62,
                                   "bus_gbs": (4 * n * 62 * 2 * (world - 1) / world / (stage_ms["exchange"] * 1e-3) / 1e9)
                                   if stage_ms.get("exchange") else None}
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N=1 only
            out["cpu_baseline"] = cpu_arm(args, steps=args.cpu_steps, warmup=0)["cpu_baseline"]
        if world == 1 and not args.no_train_iter:  # last: nothing after it needs the device
            try:
                out["train_iter"] = train_iter_section(d, dev, W, H, dev_vm[0], dev_K[0])
            except Exception as exc:  # the headline line must survive a failure of this additional section
                out["train_iter"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ------------------------------------------------------------------------------ stage-1 training iteration
def train_iter_section(d, dev, W, H, vm, K, steps=10, warmup=4):
    """ms per stage-1 training iteration (BASELINE.json metric, second half): deformation network
    (freegaussian_model.py:832-845) -> rasterization (:847-868) -> blend + L1 + SSIM loss (:875-877, 965-981) ->
    backward -> Adam step of the Gaussian groups and of the network (freegaussian_config.py:48-85).  `deform_fwd_ms` /
    `deform_bwd_ms` are CUDA-event brackets around the network's forward and around its part of `loss.backward()`.  Everything on
    the device is this repo's kernels except the network's Adam (torch, fused) and a few scalar glue ops.
    The same network in plain torch fp32 (what the reference executes) is timed beside it."""
    from freegaussian_b200.deform import DeformNetwork
    from freegaussian_b200.losses import blend_l1_ssim_loss
    from freegaussian_b200.optim import GaussianAdam
    from freegaussian_b200.rendering import rasterization

    n = d.means.shape[0]
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    net = DeformNetwork(is_blender=True).to(dev)  # freegaussian_model.py:198
    with torch.no_grad():  # small deformations, as a trained network produces: the render workload stays the scene's
        for nm in ("branch_w", "branch_v", "gaussian_rotation", "gaussian_scaling"):
            getattr(net, nm).weight.mul_(0.01)
            getattr(net, nm).bias.mul_(0.01)
    means = d.means.detach().clone().requires_grad_(True)
    scales_log = d.scales.detach().log().requires_grad_(True)        # the model stores log scales (:844)
    quats = d.quats.detach().clone().requires_grad_(True)
    op_logit = torch.logit(d.opacities.detach().clamp(1e-4, 1 - 1e-4)).requires_grad_(True)  # and logit opacities (:851)
    sh = d.sh.detach().clone().requires_grad_(True)
    gt = torch.rand(H, W, 3, generator=gen).to(dev)
    bg = torch.zeros(3, device=dev)
    t = torch.tensor([[0.3]], device=dev).expand(n, -1)
    adam = GaussianAdam.for_reference_groups(means, sh, op_logit, scales_log, quats)
    adam_net = torch.optim.Adam(net.parameters(), lr=1.6e-4 * 5, eps=1e-15, fused=True)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    last = {}

    cpu_phase = {}  # host time spent enqueueing each phase (FG_BENCH_CPU_PHASES=1 prints it): is the step launch-bound?

    def tick(name, t0):
        cpu_phase[name] = cpu_phase.get(name, 0.0) + (time.perf_counter() - t0)
        return time.perf_counter()

    def iteration(with_deform: bool):
        for p in (means, scales_log, quats, op_logit, sh):
            p.grad = None
        adam_net.zero_grad(set_to_none=True)
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        tc = time.perf_counter()
        e0.record()
        if with_deform:
            m2, s2, q2 = net.deform_gaussians(means, scales_log, quats, t)
            for out in (m2, s2, q2):  # the last of these fires when the render backward is done and the network's starts
                out.register_hook(lambda g_: e2.record())
        else:  # warm-up phase of the reference (step < warm_up, :832-833): no deformation
            m2, s2, q2 = means, torch.exp(scales_log), quats
        e1.record()
        tc = tick("deform_fwd", tc)
        render, alpha, meta = rasterization(m2, q2, s2, torch.sigmoid(op_logit), sh, vm, K, W, H, packed=False,
                                            near_plane=0.01, far_plane=1e10, render_mode="RGB+ED", sh_degree=3,
                                            sparse_grad=False, absgrad=True, rasterize_mode="classic")
        loss = blend_l1_ssim_loss(render, alpha, bg, gt, 0.2)
        tc = tick("render_fwd+loss", tc)
        loss.backward()
        e3.record()
        tc = tick("backward", tc)
        adam.step()
        if with_deform:
            adam_net.step()
        tick("adam", tc)
        last["radii"] = meta["radii"]
        return e0, e1, e2, e3

    def timed(with_deform: bool):
        for _ in range(warmup):
            iteration(with_deform)
        torch.cuda.synchronize()
        cpu_phase.clear()
        a, b = ev(), ev()
        a.record()
        marks = [iteration(with_deform) for _ in range(steps)]
        b.record()
        torch.cuda.synchronize()
        fwd = statistics.median(m[0].elapsed_time(m[1]) for m in marks)
        bwd = statistics.median(m[2].elapsed_time(m[3]) for m in marks) if with_deform else 0.0
        return a.elapsed_time(b) / steps, fwd, bwd

    l0 = None
    from freegaussian_b200 import _lib
    ms_plain, _, _ = timed(False)
    l0 = _lib.launch_count()
    ms_full, ms_deform_fwd, ms_deform_bwd = timed(True)
    if os.environ.get("FG_BENCH_CPU_PHASES"):
        print("host ms per iteration spent enqueueing:", {k: round(v * 1e3 / steps, 3) for k, v in cpu_phase.items()}, file=sys.stderr)
    launches = (_lib.launch_count() - l0) / (steps + warmup)
    n_vis = int((last["radii"] > 0).sum())

    # stage-2 step (freegaussian_control_model.py:122-179): the control network on the controllable subset (here the
    # "articulated part" of the scene: the Gaussians that move between the two frames), scattered back, same render + loss
    from freegaussian_b200.deform import ControlNetwork
    control = ControlNetwork().to(dev)
    adam_ctl = torch.optim.Adam(control.parameters(), lr=1.6e-4 * 5, eps=1e-15, fused=True)
    part = ((d.means_next - d.means).abs().sum(-1) > 0).nonzero().squeeze(1)
    value = torch.randn(part.numel(), 3, generator=gen).to(dev) * 0.05

    def iteration2():
        for p in (means, scales_log, quats, op_logit, sh):
            p.grad = None
        adam_ctl.zero_grad(set_to_none=True)
        d_xyz, d_rot, d_scale = control(means[part], value)                 # :122, :143
        m2 = means + torch.zeros_like(means).index_copy(0, part, d_xyz)      # :147-149
        s2 = torch.exp(scales_log) + torch.zeros_like(scales_log).index_copy(0, part, d_scale)   # :151-153
        q2 = quats / quats.norm(dim=-1, keepdim=True) + torch.zeros_like(quats).index_copy(0, part, d_rot)  # :155-157
        render, alpha, meta = rasterization(m2, q2, s2, torch.sigmoid(op_logit), sh, vm, K, W, H, packed=False,
                                            near_plane=0.01, far_plane=1e10, render_mode="RGB+ED", sh_degree=3,
                                            sparse_grad=False, absgrad=True, rasterize_mode="classic")
        blend_l1_ssim_loss(render, alpha, bg, gt, 0.2).backward()
        adam.step()
        adam_ctl.step()

    for _ in range(warmup):
        iteration2()
    torch.cuda.synchronize()
    a2, b2 = ev(), ev()
    a2.record()
    for _ in range(steps):
        iteration2()
    b2.record()
    torch.cuda.synchronize()
    ms_stage2 = a2.elapsed_time(b2) / steps


    # the reference's own execution of the network: torch fp32 nn.Linear / relu / cat on this GPU, forward + backward
    x = means.detach()
    lins = [net.linear[i] for i in range(8)]

    def torch_trunk():
        t_emb = net._time_row(t).expand(n, -1)
        x_emb = torch.cat([x] + [f(x * (2.0 ** k)) for k in range(10) for f in (torch.sin, torch.cos)], -1)
        h = torch.cat([x_emb, t_emb], -1)
        for i in range(8):
            h = torch.relu(torch.nn.functional.linear(h, lins[i].weight, lins[i].bias))
            if i == 4:
                h = torch.cat([x_emb, t_emb, h], -1)
        return torch.cat([torch.nn.functional.linear(h, getattr(net, nm).weight, getattr(net, nm).bias)
                          for nm in ("branch_w", "branch_v", "gaussian_rotation", "gaussian_scaling")], -1)

    tt = []
    for i in range(4):
        net.zero_grad(set_to_none=True)
        a, b, c = ev(), ev(), ev()
        a.record()
        out = torch_trunk()
        b.record()
        out.sum().backward()
        c.record()
        torch.cuda.synchronize()
        tt.append((a.elapsed_time(b), b.elapsed_time(c)))
    net.zero_grad(set_to_none=True)
    torch_fwd, torch_bwd = min(v[0] for v in tt[1:]), min(v[1] for v in tt[1:])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2  # tf32 runs at half the bf16 rate on this tensor core
    flop_fwd = 3 * 2.0 * n * (96 * 256 + 6 * 256 * 256 + 352 * 256 + 256 * 32)  # 3 tf32 products per fp32 product
    ach = flop_fwd / (ms_deform_fwd * 1e-3) / 1e12
    return {
        "ms": ms_full, "ms_without_deform": ms_plain, "deform_fwd_ms": ms_deform_fwd,
        "deform_bwd_ms": ms_deform_bwd,
        "stage2_ms": ms_stage2, "stage2_controlled_gaussians": int(part.numel()),
        "gaussians": n, "visible": n_vis, "launches_per_iter": launches,
        "config": "stage-1 step: DeformNetwork(is_blender=True) -> RGB+ED render -> blend+L1+SSIM -> backward -> Adam; "
                  "stage2_ms: ControlNetwork on the controlled subset -> the same render / loss / backward / Adam",
        "deform_roofline": {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                            "note": "forward; tf32 MMA flops issued (3xTF32) / CUDA-event time; peak = measured bf16 GEMM peak / 2"},
        "torch_fp32_network": {"fwd_ms": torch_fwd, "bwd_ms": torch_bwd,
                               "note": "the same network with torch.nn.functional.linear in fp32 on this GPU (what the reference runs)"},
    }


# ------------------------------------------------------------------------------ CPU arm
def cpu_arm(args, steps: int, warmup: int):
    """The reference algorithm (oracle restatement of the gsplat path -- gsplat itself cannot be
    installed here) on the host cores, bounded sample: the SAME scene and camera,
    all Gaussians projected, CPU_CROP x CPU_CROP centre crop of the frame composited, fwd+bwd."""
    from oracle import knn as oknn
    from oracle import render as oracle
    from freegaussian_b200.scenes import make_scene

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, W, H = WORKLOADS[args.workload]
    if torch.cuda.is_available():
        from freegaussian_b200.knn import k_nearest
        knn3 = lambda m: k_nearest(m.cuda(), 3)[0].cpu()  # scene construction only (untimed)
    else:
        knn3 = lambda m: torch.from_numpy(oknn.reference_knn(m.numpy(), 3)[0])
    sc = make_scene(n, W, H, n_views=N_VIEW_POOL, recipe=args.recipe, seed=0, knn3=knn3)
    cw, ch = min(CPU_CROP, W), min(CPU_CROP, H)
    K = sc.Ks[:1].clone()
    K[:, 0, 2] -= (W - cw) / 2
    K[:, 1, 2] -= (H - ch) / 2
    w_rgbd, w_flow = loss_weights(ch, cw, 1)
    params = [sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.means_next]
    for p in params:
        p.requires_grad_(True)
    times = []
    for i in range(warmup + steps):
        for p in params:
            p.grad = None
        t0 = time.perf_counter()
        r, a, m = oracle.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.sh, sc.viewmats[:1], K, cw, ch,
                                       near_plane=0.01, far_plane=1e10, render_mode="RGB+ED", sh_degree=3,
                                       means_next=sc.means_next)
        loss = (r * w_rgbd).sum() + (m["flow"] * w_flow).sum()
        loss.backward()
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    t = sum(times) / len(times)
    val = cw * ch / t / 1e6
    sample = (f"{steps} step(s) of fwd+bwd on a {cw}x{ch} centre crop of the {W}x{H} frame, all {n} Gaussians "
              f"projected, {m['flatten_ids'].numel()} intersections in the crop; {t:.2f} s/step")
    base = {"value": val, "unit": "MPix/s", "cores": cores, "kind": "port", "sample": sample}
    return {"cpu_baseline": base, "ms_per_step": t * 1e3, "config_workload": f"{args.workload}: {n} Gaussians, {W}x{H}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_arm(args, steps=args.steps, warmup=min(args.warmup, 1))
    n, W, H = WORKLOADS[args.workload]
    val = res["cpu_baseline"]["value"]
    out = {
        "impl": "reference",
        "metric": "fwd+bwd MPix/s (RGB+depth+flow) at 1M Gaussians",
        "value": val, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n} Gaussians, {W}x{H}, 1 view/step, SH3, RGB+ED+flow (6 ch), "
                               f"{args.recipe} scene", "note": "CPU arm: bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": res["cpu_baseline"],
        "e2e": {"value": val, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
