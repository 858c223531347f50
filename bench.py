#!/usr/bin/env python
"""bench.py -- fwd+bwd MPix/s (RGB+depth+flow) of the splat-render hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (config.workload = "cfg3"): 1 M Gaussians, 1920x1080, SH degree 3, render_mode
"RGB+ED" + rendered flow (6 composited channels), one view per GPU per step (view-sharded,
weak scaling), scalar loss = <render, w_rgbd> + <flow, w_flow> (gradient planes w_rgbd / w_flow), backward to all
Gaussian parameters; at N>1 the exchange of the parameter gradients runs inside the backward
(freegaussian_b200/dist.py::ViewShardedExchange, csrc/exchange.cu).  Synthetic "trained-like" scene
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
CPU_CROP = 128   # the CPU arm composites a CPU_CROP x CPU_CROP centre crop of the same frame
METRIC = "fwd+bwd MPix/s (RGB+depth+flow) at 1M Gaussians"


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
    ap.add_argument("--no-extras", action="store_true", help="skip the knn / configs sections")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = published colour gradients + in-switch all-reduce inside the backward (csrc/exchange.cu); "
                         "'nccl' = one NCCL all-reduce over the dense 236 B/Gaussian arena (round-1 path, kept for comparison)")
    return ap.parse_args()


def workload_string(workload: str, recipe: str, views: int = 1) -> str:
    """config.workload -- the SAME string on both arms (the driver compares it)."""
    n, W, H = WORKLOADS[workload]
    return (f"{workload}: {n} Gaussians, {W}x{H}, {views} view/GPU/step, SH3, RGB+ED+flow (6 ch), {recipe} scene")


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


def quiesce_gc() -> None:
    """Before a timed region: collect once, then move everything alive (torch / numpy / sklearn module state, the scene) to the
    permanent generation, so that a generation-2 collection walking ~10^6 long-lived objects (tens of ms of host time, seen as
    one 37 ms iteration among thirty 10 ms ones) cannot land inside the region.  Garbage produced inside the region is still
    collected."""
    import gc
    gc.collect()
    gc.freeze()


def bind_to_gpu_numa_node(local_rank: int) -> None:
    """Pin this rank's host threads to the CPUs next to its GPU (nvidia-smi topo's "CPU Affinity"), so the pinned staging
    buffers are first-touched on the GPU's NUMA node: at 8 ranks the per-step host->device copies otherwise cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# ------------------------------------------------------------------------------ workload
def build_scene(workload: str, recipe: str, device, n_views: int):
    from freegaussian_b200.knn import k_nearest
    from freegaussian_b200.scenes import make_scene
    n, w, h = WORKLOADS[workload]
    knn3 = lambda m: k_nearest(m.to(device), 3)[0].cpu()  # the product KNN kernel seeds the scales (model.py:158)
    return make_scene(n, w, h, n_views=n_views, recipe=recipe, seed=0, knn3=knn3)


def loss_weights(h: int, w: int, seed: int):
    """The per-step "ground truth" planes of the scalar loss, in the formats a dataloader holds them in on the host:
    rgb uint8 [h,w,3] (nerfstudio caches images as uint8), depth float16 [h,w,1] (16-bit depth maps), flow float16 [h,w,2]."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randint(0, 256, (1, h, w, 3), generator=g, dtype=torch.uint8)
    depth = torch.rand(1, h, w, 1, generator=g).to(torch.float16)
    flow = (torch.rand(1, h, w, 2, generator=g) * 0.1).to(torch.float16)
    return rgb, depth, flow


def decode_weights(rgb_u8, depth, flow_h):
    """uint8 / fp16 host formats -> the float32 weight images of the loss (device-side, as a trainer decodes a batch)."""
    w_rgbd = torch.cat([rgb_u8.to(torch.float32) * (1.0 / 255.0), depth.to(torch.float32)], -1)
    return w_rgbd, flow_h.to(torch.float32)


class Job:
    """One workload on this rank: scene, parameters, host-side per-step inputs, and the step itself."""

    def __init__(self, args, workload, recipe, V, dev, rank, world, xchg, view_pool=N_VIEW_POOL):
        from freegaussian_b200.dist import DensificationStats
        self.args, self.workload, self.recipe, self.V = args, workload, recipe, V
        self.dev, self.rank, self.world, self.xchg = dev, rank, world, xchg
        n, W, H = WORKLOADS[workload]
        self.n, self.W, self.H = n, W, H
        self.sc = build_scene(workload, recipe, dev, max(view_pool, V) * world)
        self.d = self.sc.to(dev)
        d = self.d
        self.params = [d.means, d.quats, d.scales, d.opacities, d.sh, d.means_next]
        for p in self.params:
            p.requires_grad_(True)
        sc = self.sc
        self.my_views = list(range(rank, max(view_pool, V) * world, world))  # dist.shard_views
        mv = self.my_views
        pick = lambda t, j: torch.cat([t[mv[(j + k) % len(mv)]][None] for k in range(V)], 0)  # noqa: E731
        self.host_vm = [pick(sc.viewmats, j).clone().pin_memory() for j in range(len(mv))]
        self.host_K = [pick(sc.Ks, j).clone().pin_memory() for j in range(len(mv))]
        rgb, depth, flow = loss_weights(H, W, 1 + rank)
        self.host_w = [t.repeat(V, 1, 1, 1).pin_memory() for t in (rgb, depth, flow)]
        self.w_rgbd, self.w_flow = decode_weights(*[t.to(dev) for t in self.host_w])
        self.dev_vm = [v.to(dev) for v in self.host_vm]
        self.dev_K = [k.to(dev) for k in self.host_K]
        self.stats = DensificationStats(n, dev)
        self.info = {}
        self.it = 0
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staged = {}
        self.loss_ring = torch.full((64,), float("nan")).pin_memory()
        self.quantiles = {}
        # two device-side staging sets, filled alternately on the copy stream (a dataloader's double buffer): the raw
        # uint8 / f16 planes as copied, and the float32 weight images they are decoded into on the same stream
        self.stage_bufs = [tuple(torch.empty_like(t, device=dev) for t in (self.host_vm[0], self.host_K[0], *self.host_w))
                           for _ in range(2)]
        self.decoded = [(self.w_rgbd.clone(), self.w_flow.clone()) for _ in range(2)]
        self.consumed = [None, None]  # event: the step that read staging set k has finished with it
        self.h2d_bytes = int(sum(t.numel() * t.element_size() for t in (self.host_vm[0], self.host_K[0], *self.host_w)))

    def stage_inputs(self, i: int):
        """H2D copy of step i's host inputs (pinned) on the copy stream: the usual input prefetch --
        step i+1's camera and target images travel while step i computes, all inside the timed region."""
        j, k = i % len(self.my_views), i % 2
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.copy_stream.wait_event(self.consumed[k])
            bufs = self.stage_bufs[k]
            for dst, src in zip(bufs, (self.host_vm[j], self.host_K[j], *self.host_w)):
                dst.copy_(src, non_blocking=True)
            # decode behind the copy, still on the copy stream (what a data-loading stream does): uint8 -> [0,1], f16 -> f32
            wr, wf = self.decoded[k]
            torch.mul(bufs[2], 1.0 / 255.0, out=wr[..., :3])
            wr[..., 3:].copy_(bufs[3])
            wf.copy_(bufs[4])
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.staged[i] = ((bufs[0], bufs[1], wr, wf), ev)

    def step(self, i: int, e2e: bool):
        from freegaussian_b200 import rendering
        from freegaussian_b200.dist import exchange
        from freegaussian_b200.rendering import rasterization
        d, W, H = self.d, self.W, self.H
        j = i % len(self.my_views)
        if e2e:
            if i not in self.staged:
                self.stage_inputs(i)
            (vm, K, wr, wf), ev = self.staged.pop(i)
            torch.cuda.current_stream().wait_event(ev)
            self.stage_inputs(i + 1)  # prefetch (copy + decode) the next step's inputs behind this step's kernels
        else:
            vm, K, wr, wf = self.dev_vm[j], self.dev_K[j], self.w_rgbd, self.w_flow
        for p in self.params:
            p.grad = None
        render, alpha, meta = rasterization(d.means, d.quats, d.scales, d.opacities, d.sh, vm, K, W, H,
                                            packed=False, near_plane=0.01, far_plane=1e10, render_mode="RGB+ED",
                                            sh_degree=3, sparse_grad=False, absgrad=True, rasterize_mode="classic",
                                            means_next=d.means_next)
        meta["means2d"].retain_grad()
        # weighted sums as dot products: one reduction kernel each, no 33 MB product tensor in between (the loss only
        # exists to hand the compositing backward dense, non-trivial upstream gradients: v_render = wr, v_flow = wf)
        # at N>1 with --exchange peer the cross-rank sum of every parameter gradient happens INSIDE this backward
        if render.shape[0] == 1:
            # the loss is linear in the images, so its gradient with respect to them IS (w_rgbd, w_flow): the planes go to
            # autograd as the upstream gradients directly (what `loss.backward()` computes, minus two 33 MB multiplications
            # by the scalar 1.0), and the loss value itself is two dot products outside the graph
            with torch.no_grad():
                loss = torch.dot(render.reshape(-1), wr.reshape(-1)) + torch.dot(meta["flow"].reshape(-1), wf.reshape(-1))
            torch.autograd.backward([render, meta["flow"]], [wr.view_as(render), wf.view_as(meta["flow"])])
        else:  # several views per step share the planes
            loss = (render * wr).sum() + (meta["flow"] * wf).sum()
            loss.backward()
        self.stats.accumulate_local(meta["radii"], meta["means2d"].absgrad, H, W)
        if self.xchg is None and self.world > 1:  # --exchange nccl: one all-reduce over the dense arena
            with rendering._stage("exchange"):
                exchange([p.grad for p in self.params])
        # The densification statistics are accumulated per rank by one kernel per step and reduced across ranks when
        # they are consumed (refine_every = 100 steps in the reference configs) -- same numbers.
        self.it += 1
        if self.it % REFINE_EVERY == 0:
            self.stats.sync()
        self.info["meta"] = meta
        if e2e:
            self.consumed[i % 2] = torch.cuda.Event()
            self.consumed[i % 2].record()
            # D2H read of the step's result: 4 bytes into a pinned ring, read by the host once the copy has
            # landed (checked at the next step's list-size sync and at the end of the timed region), the way a
            # training loop logs its loss without stalling the launch queue
            self.loss_ring[i % len(self.loss_ring)].copy_(loss.detach(), non_blocking=True)

    def timed(self, k: int, e2e: bool) -> float:
        import torch.distributed as dist
        world, dev = self.world, self.dev
        quiesce_gc()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        self.loss_ring.fill_(float("nan"))
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        marks[0].record()
        for i in range(k):
            self.step(i, e2e)
            marks[i + 1].record()
        torch.cuda.synchronize()
        if e2e:
            assert bool(torch.isfinite(self.loss_ring[:min(k, len(self.loss_ring))]).all()), "a step's loss never reached the host"
        per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(k))
        pct = lambda q: per_step[min(k - 1, int(q * k))]  # noqa: E731
        self.quantiles["e2e" if e2e else "dev"] = {"p10": pct(0.10), "p50": pct(0.50), "p90": pct(0.90)}
        if world > 1:
            dist.barrier()
        ms = torch.tensor([marks[0].elapsed_time(marks[-1])], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(self, steps: int, warmup: int, e2e: bool = True):
        from freegaussian_b200 import _lib
        # priming: one untimed step per camera of this rank's pool, so that every buffer whose size depends on the view (tile
        # lists, coarse pairs) has reached its capacity before anything is timed -- a view first met inside the timed region
        # costs its rank a cudaMalloc + a second list-building call, and at N > 1 every other rank waits for it in the exchange
        for i in range(len(self.my_views)):
            self.step(i, False)
        for i in range(max(warmup, 3)):
            self.step(i, False)
        l0 = _lib.launch_count()
        ms_dev = self.timed(steps, False)
        launches = _lib.launch_count() - l0
        ms_e2e = None
        if e2e:
            self.step(0, True)
            self.staged.clear()
            ms_e2e = self.timed(steps, True)
            self.staged.clear()
        return ms_dev, ms_e2e, launches


def run_ours(args):
    import torch.distributed as dist
    from freegaussian_b200 import _lib
    from freegaussian_b200 import rendering
    from freegaussian_b200.dist import ViewShardedExchange

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
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
    xchg, xchg_note = None, None
    if world > 1 and args.exchange == "peer":
        # the exchange runs inside the projection backward over NVLink peer memory (csrc/exchange.cu); a box without peer
        # access / symmetric memory falls back -- on every rank together -- to the NCCL all-reduce of the dense arena
        ok = 1
        try:
            xchg = ViewShardedExchange()
            xchg.prepare(1 << 20, 0, 0, dev)
        except Exception as exc:  # noqa: BLE001
            ok, xchg_note = 0, f"{type(exc).__name__}: {exc}"
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            xchg.install()
        else:
            xchg, xchg_note = None, xchg_note or "symmetric memory unavailable on another rank"

    V = args.views_per_gpu
    job = Job(args, args.workload, args.recipe, V, dev, rank, world, xchg)
    n, W, H = job.n, job.W, job.H

    # nvidia-smi takes a few hundred ms to deliver its first sample, longer than a timed region: it is started before
    # the warm-up and stopped after the second timed region, so its samples cover warm-up + both timed regions (all under load)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ms_e2e, launches = job.measure(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing + roofline (every rank runs the steps -- they contain collectives --
    #      rank 0 records the CUDA-event time of each C-ABI stage)
    out = None
    rendering.stage_timer.enabled = rank == 0
    rendering.stage_timer.reset()
    for i in range(min(args.steps, 8)):
        job.step(i, False)
    torch.cuda.synchronize()
    rendering.stage_timer.enabled = False
    if rank == 0:
        stage_ms = {k: statistics.mean(v) for k, v in rendering.stage_timer.summary().items()}
        meta = job.info["meta"]
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
        CN = V * n
        if rendering.SORT_MODE == "key64":
            cam_bits, tile_bits = 1, int(math.floor(math.log2(n_tiles))) + 1
            passes = (32 + tile_bits + cam_bits + 7) // 8
            sort_bytes, emit_bytes = M * (8 + passes * 24), CN * 20 + M * 12
        else:  # two-level: 32-bit tile keys, ceil(log2(tiles)/8) passes; depth sort of the n splats separately
            passes = (max(1, math.ceil(math.log2(n_tiles))) + 7) // 8
            sort_bytes, emit_bytes = M * (4 + passes * 16), CN * 24 + M * 8
        algo = {  # algorithmic bytes / flops per launch (DESIGN.md "Kernels")
            "project_fwd": ("hbm", n_vis * 276 + (CN - n_vis) * 44),
            "project_bwd": ("hbm", n_vis * 548 + (CN - n_vis) * (44 + 4 + 236)),
            "sort": ("hbm", sort_bytes),
            "depth_sort": ("hbm", CN * (8 + 8) + CN * (4 + 4 * 16)),
            "emit": ("hbm", emit_bytes),
            "bin_count": ("hbm", CN * 4 + n_vis * 16 + CN * 12),
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
        # captures (profiles/); only valid for the configuration they were taken on
        ncu_traffic = {}
        if args.workload == "cfg3" and args.recipe == "trained_like" and V == 1:
            ncu_traffic = NCU_TRAFFIC
        for name, k in kernels.items():
            k["traffic"] = ncu_traffic.get(name)
        comp = {k: v for k, v in stage_ms.items() if not k.startswith("xchg_")}
        dominant = max(comp, key=comp.get)
        roof = dict(kernels.get(dominant, {}))
        roof.update({"kernel": dominant, "traffic": ncu_traffic.get(dominant),
                     "peak_source": "fg_measure_fp32_tflops (FFMA microbenchmark, this run)" if roof.get("bound") == "fp32" else hbm_src})
        hbm_kernels = {k: v for k, v in kernels.items() if v["bound"] == "hbm"}
        dom_hbm = max(hbm_kernels, key=lambda k: hbm_kernels[k]["ms"]) if hbm_kernels else None

        pix = world * V * W * H
        out = {
            "metric": METRIC,
            "value": pix * args.steps / (ms_dev * 1e-3) / 1e6,
            "unit": "MPix/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps,
            "ms_per_step_quantiles": job.quantiles.get("dev"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, args.recipe, V), "views_per_step": world * V,
                       "l2": "inputs larger than L2 (no flush)", "priming_steps": len(job.my_views),
                       "n_isects": M, "visible": n_vis, "sort_mode": rendering.SORT_MODE, "pairs_per_pixel": pairs / (V * W * H),
                       "parallelism": f"view-sharded dp{world}" if world > 1 else "single GPU"},
            "e2e": {"value": pix * args.steps / (ms_e2e * 1e-3) / 1e6, "unit": "MPix/s",
                    "h2d_bytes_per_step": job.h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "ms_per_step_quantiles": job.quantiles.get("e2e"),
                    "host_formats": "cameras f32; target planes as a dataloader holds them: rgb uint8, depth f16, flow f16; copied "
                                    "from pinned memory and decoded to f32 on a copy stream, all inside the timed region"},
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
                ar = stage_ms.get("xchg_allreduce")
                ex = stage_ms.get("exchange", 0.0) - stage_ms.get("project_bwd_geo", 0.0)  # the stage brackets the geometry kernel too
                out["exchange"] = {"mode": "peer: published colour gradients (12 B per visible (view, Gaussian)) read over NVLink by "
                                           "fg_xchg_sh_bwd_views + in-switch two-shot all-reduce of the geometry gradients",
                                   "multicast": bool(xchg.multicast), "ms": ex,
                                   "allreduce_ms": ar, "sh_views_ms": stage_ms.get("xchg_sh_views"),
                                   "overlap": "fg_xchg_sh_bwd_views runs on a side stream under the geometry kernel and the all-reduce",
                                   "allreduce_bytes": geo_b,
                                   "allreduce_bus_gbs": geo_b * 2 * (world - 1) / world / (ar * 1e-3) / 1e9 if ar else None,
                                   "published_bytes_per_rank": n_vis * 12 + V * n // 8,
                                   "dense_arena_bytes_replaced": 4 * n * 62}
            else:
                out["exchange"] = {"mode": "nccl: one all-reduce over the dense gradient arena", "fallback_reason": xchg_note,
                                   "ms": stage_ms.get("exchange"),
                                   "allreduce_bytes": 4 * n * 62,
                                   "bus_gbs": (4 * n * 62 * 2 * (world - 1) / world / (stage_ms["exchange"] * 1e-3) / 1e9)
                                   if stage_ms.get("exchange") else None}
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N=1 only
            out["cpu_baseline"] = cpu_arm(args, steps=args.cpu_steps, warmup=0, scene=job.sc)["cpu_baseline"]

    # ---- other BASELINE.json configs (ms/step only; parity for them is in tests/test_gpu_baseline_shapes.py)
    if not args.no_extras and args.workload == "cfg3" and V == 1:
        extra = {}
        try:
            if world == 1:
                del job
                torch.cuda.empty_cache()
                for name in ("cfg1", "cfg2"):
                    vv = 4 if name == "cfg1" else 1
                    j2 = Job(args, name, args.recipe, vv, dev, rank, world, xchg, view_pool=4)
                    ms2, _, l2 = j2.measure(20, 5, e2e=False)
                    extra[name] = {"workload": workload_string(name, args.recipe, vv), "ms_per_step": ms2 / 20,
                                   "MPix/s": vv * j2.W * j2.H * 20 / (ms2 * 1e-3) / 1e6, "gpu_launches_per_step": l2 / 20,
                                   "quantiles": j2.quantiles.get("dev")}
                    del j2
                extra["knn"] = knn_section(dev)
            elif world == 8:
                del job
                torch.cuda.empty_cache()
                rendering.stage_timer.reset()
                j4 = Job(args, "cfg4", args.recipe, 4, dev, rank, world, xchg, view_pool=4)
                ms4, _, _ = j4.measure(6, 3, e2e=False)
                rendering.stage_timer.enabled = rank == 0
                rendering.stage_timer.reset()
                for i in range(3):
                    j4.step(i, False)
                torch.cuda.synchronize()
                rendering.stage_timer.enabled = False
                st4 = {k: statistics.mean(v) for k, v in rendering.stage_timer.summary().items()} if rank == 0 else {}
                extra["cfg4"] = {"workload": workload_string("cfg4", args.recipe, 4), "views_per_step": 32, "ms_per_step": ms4 / 6,
                                 "MPix/s": 32 * j4.W * j4.H * 6 / (ms4 * 1e-3) / 1e6, "exchange_ms": st4.get("exchange"),
                                 "allreduce_ms": st4.get("xchg_allreduce"), "stage_ms": st4}
                del j4
        except Exception as exc:  # the headline line must survive a failure of an additional section
            extra["error"] = f"{type(exc).__name__}: {exc}"
        if rank == 0:
            out["configs"] = extra
    if rank == 0 and world == 1 and not args.no_train_iter:  # last: nothing after it needs the device
        try:
            sc = build_scene(args.workload, args.recipe, dev, 1)
            d = sc.to(dev)
            out["train_iter"] = train_iter_section(d, dev, W, H, d.viewmats[:1].contiguous(), d.Ks[:1].contiguous())
        except Exception as exc:
            out["train_iter"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            os._exit(0)


# dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full` captures under profiles/ (cfg3, trained_like)
NCU_TRAFFIC = {"rasterize_bwd": 127.1e6, "rasterize_fwd": 25.9e6, "project_bwd": 403.5e6, "project_fwd": 148.0e6,
               "fine_bin": 103.0e6}


# ------------------------------------------------------------------------------ k-NN (BASELINE cfg5)
def knn_section(dev, n=3_000_000, k=16, sample=100_000):
    """`knn_gaussian` / init k-NN at the cfg5 shape: k=16 over 3 M points (freegaussian_model.py:293-311).  GPU: CUDA events
    around fg_knn_f32 (grid build + query).  CPU: the reference's sklearn call, fit on all points + query of a bounded
    sample, on 1 core and with n_jobs=-1; the kernel's distances on that sample must equal sklearn's bit for bit."""
    import numpy as np
    from sklearn.neighbors import NearestNeighbors

    from freegaussian_b200.knn import k_nearest
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(n, 3, generator=g) - 0.5) * 6.0  # freegaussian_model.py:155 recipe
    xd = x.to(dev)
    k_nearest(xd, k)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist_gpu, idx_gpu = k_nearest(xd, k)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    xn = x.numpy()
    q = xn[:sample]
    out = {}
    for label, jobs in (("sklearn_1core", 1), ("sklearn_all_cores", -1)):
        t0 = time.perf_counter()
        model = NearestNeighbors(n_neighbors=k + 1, algorithm="auto", metric="euclidean", n_jobs=jobs).fit(xn)
        t1 = time.perf_counter()
        dref, iref = model.kneighbors(q)
        t2 = time.perf_counter()
        out[label] = {"fit_s": t1 - t0, "query_s_sample": t2 - t1, "query_s_extrapolated": (t2 - t1) * n / sample,
                      "queries_per_s": sample / (t2 - t1)}
    same = bool(np.array_equal(dist_gpu[:sample].cpu().numpy(), dref[:, 1:].astype(np.float32)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    ppc = max(2.0, 0.5 * (k + 1))  # points per grid cell the kernel sizes its cells for (csrc/knn.cu)
    stream_bytes = n * 27 * ppc * 12   # 12 B x points in the 27 neighbour cells, per query
    floor_bytes = n * 12 + n * k * 8   # compulsory: read the points once, write distances + indices
    return {"points": n, "k": k, "ms": ms, "queries_per_s": n / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": stream_bytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": stream_bytes / (ms * 1e-3) / 1e9 / hbm,
                         "note": "candidate streaming, 12 B x 27 cells x points per cell per query (SURVEY 8(d)); these bytes "
                                 "are served by L1/L2 (points are cell-sorted), the kernel is bound by the float64 top-k insertion",
                         "compulsory_gbs": floor_bytes / (ms * 1e-3) / 1e9},
            "cpu": dict(out, cores=os.cpu_count(), sample=f"fit on all {n} points, query of the first {sample}"),
            "speedup_vs_sklearn_1core": (out["sklearn_1core"]["fit_s"] + out["sklearn_1core"]["query_s_extrapolated"]) / (ms * 1e-3),
            "bit_exact_distances_on_sample": same}


# ------------------------------------------------------------------------------ stage-1 training iteration
def train_iter_section(d, dev, W, H, vm, K, steps=30, warmup=20):
    """ms per stage-1 training iteration (BASELINE.json metric, second half): deformation network
    (freegaussian_model.py:832-845) -> rasterization (:847-868) -> blend + L1 + SSIM loss (:875-877, 965-981) ->
    backward -> Adam step of the Gaussian groups and of the network (freegaussian_config.py:48-85).  `deform_fwd_ms` /
    `deform_bwd_ms` are CUDA-event brackets around the network's forward and around its part of `loss.backward()`.  Everything on
    the device is this repo's kernels except the network's Adam (torch.optim.Adam(fused=True), 0.6 M weights), the time
    branch (one row) and a few scalar glue ops.  Per-iteration times are CUDA-event brackets (first event of an iteration to
    the first event of the next: every gap included); `ms` = mean of the steady-state iterations, p10 / p50 / p90 / max over the
    same, after `warmup` untimed iterations; the first iteration after the synchronize is reported apart.  The same network in plain torch fp32 (what the reference executes) is timed beside it."""
    from freegaussian_b200 import _lib
    from freegaussian_b200.deform import ControlNetwork, DeformNetwork
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
    # ground truth = the scene's own render with perturbed colours: a training state near its optimum, so the workload is
    # stationary.  (A noise image as target makes the optimiser blur the scene -- every Gaussian inflates, the tile lists grow
    # from 33 M to 109 M entries within 50 iterations, r2 measurement -- and the "iteration" timed is a different one each time.)
    with torch.no_grad():
        sh_gt = d.sh.detach() + 0.05 * torch.randn(d.sh.shape, generator=gen).to(dev)
        gt = rasterization(d.means, d.quats, d.scales, d.opacities, sh_gt, vm, K, W, H, packed=False, near_plane=0.01,
                           far_plane=1e10, render_mode="RGB", sh_degree=3, rasterize_mode="classic")[0][0].clamp(0, 1).contiguous()
        del sh_gt
    bg = torch.zeros(3, device=dev)
    t = torch.tensor([[0.3]], device=dev).expand(n, -1)
    adam = GaussianAdam.for_reference_groups(means, sh, op_logit, scales_log, quats)
    adam_net = torch.optim.Adam(net.parameters(), lr=1.6e-4 * 5, eps=1e-15, fused=True)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    last = {}
    cpu_phase = {}  # host time spent enqueueing each phase: is the step launch-bound?
    diag = {}       # allocator / host diagnostics of the timed regions

    phase_log = []  # per iteration: host ms of each phase (the slowest iteration is reported)

    def tick(name, t0):
        dt = time.perf_counter() - t0
        cpu_phase[name] = cpu_phase.get(name, 0.0) + dt
        if name == "deform_fwd":
            phase_log.append({})
        phase_log[-1][name] = round(dt * 1e3, 3)
        return time.perf_counter()

    def iteration(with_deform: bool):
        for p in (means, scales_log, quats, op_logit, sh):
            p.grad = None
        adam_net.zero_grad(set_to_none=True)
        e0, e1, e2, e3, e4 = ev(), ev(), ev(), ev(), ev()
        tc = time.perf_counter()
        e0.record()
        if with_deform:
            m2, s2, q2 = net.deform_gaussians(means, scales_log, quats, t)
            for o_ in (m2, s2, q2):  # the last of these fires when the render backward is done and the network's starts
                o_.register_hook(lambda g_: e2.record())
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
        e4.record()
        tick("adam", tc)
        last["radii"] = meta["radii"]
        return e0, e1, e2, e3, e4

    def timed(with_deform: bool):
        for _ in range(warmup):
            iteration(with_deform)
        torch.cuda.synchronize()
        quiesce_gc()
        cpu_phase.clear()
        mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        from freegaussian_b200 import rendering as _R
        _R.STATS.update(list_len_min=0, list_len_max=0)
        lists0 = dict(_R.STATS)
        a, b = ev(), ev()
        a.record()
        host_t = []
        marks = []
        for _ in range(steps):
            t_h = time.perf_counter()
            marks.append(iteration(with_deform))
            host_t.append((time.perf_counter() - t_h) * 1e3)
        b.record()
        torch.cuda.synchronize()
        diag["cuda_mallocs_in_timed_region"] = diag.get("cuda_mallocs_in_timed_region", 0) + (
            torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0)
        for k_ in ("list_capacity_changes", "list_second_call"):
            diag[k_] = diag.get(k_, 0) + _R.STATS[k_] - lists0[k_]
        diag.setdefault("tile_list_lengths", []).append([_R.STATS["list_len_min"], _R.STATS["list_len_max"]])
        if max(host_t) > diag.get("host_ms_per_iteration_max", 0.0):
            k = host_t.index(max(host_t))
            m = marks[k]
            diag["host_ms_per_iteration_max"] = max(host_t)
            diag["slowest_iteration_host_phases_ms"] = phase_log[len(phase_log) - steps + k]
            diag["slowest_iteration_device_phases_ms"] = {
                "deform_fwd": m[0].elapsed_time(m[1]), "render_to_end_of_backward": m[1].elapsed_time(m[3]),
                "adam": m[3].elapsed_time(m[4]), "index": k}
        fwd = statistics.median(m[0].elapsed_time(m[1]) for m in marks)
        bwd = statistics.median(m[2].elapsed_time(m[3]) for m in marks) if with_deform else 0.0
        # iteration i: from its first event to the first event of iteration i+1 (back-to-back, includes every gap)
        raw = [marks[i][0].elapsed_time(marks[i + 1][0]) for i in range(steps - 1)]
        if os.environ.get("FG_BENCH_PHASES"):
            for i in (0, 1, steps // 2):
                m = marks[i]
                print(f"iteration {i}: deform_fwd {m[0].elapsed_time(m[1]):.2f} render+loss+render_bwd {m[1].elapsed_time(m[2]) if with_deform else -1:.2f} "
                      f"net_bwd {m[2].elapsed_time(m[3]) if with_deform else -1:.2f} adam {m[3].elapsed_time(m[4]):.2f} total {raw[i]:.2f}", file=sys.stderr)
        # The first iteration after the synchronize starts from a drained queue (and now and then absorbs a host-side
        # hiccup of tens of ms: seen at iteration 0 only); `ms` is the mean of the steady-state iterations 1 .. K-2, the
        # first one is reported separately.
        steady = raw[1:]
        per = sorted(steady)
        q = lambda f: per[min(len(per) - 1, int(f * len(per)))]  # noqa: E731
        return (sum(steady) / len(steady), fwd, bwd,
                {"p10": q(0.1), "p50": q(0.5), "p90": q(0.9), "max": per[-1], "first_after_sync": raw[0],
                 "mean_including_first": a.elapsed_time(b) / steps})

    ms_plain, _, _, q_plain = timed(False)
    l0 = _lib.launch_count()
    ms_full, ms_deform_fwd, ms_deform_bwd, q_full = timed(True)
    if os.environ.get("FG_BENCH_TIMELINE") == "hunt":  # 40 profiled iterations: the longest kernels and the largest idle gaps
        from torch.profiler import ProfilerActivity, profile
        try:
            # which allocation goes to cudaMalloc: the allocator's own event trace, with Python stacks
            torch.cuda.memory._record_memory_history(max_entries=200000, context="alloc", stacks="python")
            for _ in range(40):
                iteration(True)
            torch.cuda.synchronize()
            snap = torch.cuda.memory._snapshot()
            torch.cuda.memory._record_memory_history(enabled=None)
            print("# cudaMalloc calls of 40 iterations (size MiB, innermost repo frames):", file=sys.stderr)
            for trace in snap.get("device_traces", []):
                for k, e in enumerate(trace):
                    if e.get("action") != "segment_alloc":
                        continue
                    # the segment is created for the allocation that follows it in the trace
                    frames = e.get("frames") or (trace[k + 1].get("frames") if k + 1 < len(trace) else []) or []
                    mine = [f"{os.path.basename(f['filename'])}:{f['line']}" for f in frames
                            if "/site-packages/" not in f.get("filename", "") and f.get("filename", "").endswith(".py")][:4]
                    print(f"{e['size'] / 2**20:10.1f}  {' < '.join(mine)}", file=sys.stderr)
        except Exception as exc:  # a diagnostic: never in the way of the run
            print(f'# allocator trace unavailable: {exc!r}', file=sys.stderr)
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(40):
                iteration(True)
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        longest = sorted(evs, key=lambda e: e.time_range.end - e.time_range.start, reverse=True)[:6]
        print("# longest device activities (us):", file=sys.stderr)
        for e in longest:
            print(f"{e.time_range.end - e.time_range.start:10.1f}  {e.name[:110]}", file=sys.stderr)
        gaps, t_prev, prev = [], evs[0].time_range.end, evs[0]
        for e in evs[1:]:
            gaps.append((e.time_range.start - t_prev, prev.name[:60], e.name[:60]))
            if e.time_range.end > t_prev:
                t_prev, prev = e.time_range.end, e
        print("# largest idle gaps (us): gap, after, before", file=sys.stderr)
        for g in sorted(gaps, reverse=True)[:6]:
            print(f"{g[0]:10.1f}  after {g[1]}  | before {g[2]}", file=sys.stderr)
        cpu = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU),
                     key=lambda e: e.time_range.end - e.time_range.start, reverse=True)[:8]
        print("# longest host-side ops (us):", file=sys.stderr)
        for e in cpu:
            print(f"{e.time_range.end - e.time_range.start:10.1f}  {e.name[:110]}", file=sys.stderr)
    elif os.environ.get("FG_BENCH_TIMELINE"):  # device timeline of one iteration (kernel, duration, idle gap before it) on stderr
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                iteration(True)
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        starts = [i for i, e in enumerate(evs) if "deform_embed" in e.name]
        if len(starts) >= 2:
            one = evs[starts[-2]:starts[-1]]
            t_prev, busy = one[0].time_range.start, 0.0
            print("# device timeline of one training iteration (us): gap-before  duration  kernel", file=sys.stderr)
            for e in one:
                gap, dur = e.time_range.start - t_prev, e.time_range.end - e.time_range.start
                busy += dur
                if gap > 5 or dur > 50:
                    print(f"{gap:9.1f} {dur:9.1f}  {e.name[:100]}", file=sys.stderr)
                t_prev = max(t_prev, e.time_range.end)
            span = t_prev - one[0].time_range.start
            print(f"# {len(one)} kernels, span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us", file=sys.stderr)
    host_ms = {k: round(v * 1e3 / steps, 3) for k, v in cpu_phase.items()}
    launches = (_lib.launch_count() - l0) / (steps + warmup)
    n_vis = int((last["radii"] > 0).sum())

    # stage-2 step (freegaussian_control_model.py:122-179): the control network on the controllable subset (here the
    # "articulated part" of the scene: the Gaussians that move between the two frames), scattered back, same render + loss
    control = ControlNetwork().to(dev)
    adam_ctl = torch.optim.Adam(control.parameters(), lr=1.6e-4 * 5, eps=1e-15, fused=True)
    part = ((d.means_next - d.means).abs().sum(-1) > 0).nonzero().squeeze(1)
    value = torch.randn(part.numel(), 3, generator=gen).to(dev) * 0.05

    def iteration2():
        for p in (means, scales_log, quats, op_logit, sh):
            p.grad = None
        adam_ctl.zero_grad(set_to_none=True)
        e0 = ev()
        e0.record()
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
        return e0

    for _ in range(warmup):
        iteration2()
    torch.cuda.synchronize()
    quiesce_gc()
    a2, b2 = ev(), ev()
    a2.record()
    marks2 = [iteration2() for _ in range(steps)]
    b2.record()
    torch.cuda.synchronize()
    raw2 = [marks2[i].elapsed_time(marks2[i + 1]) for i in range(steps - 1)]
    ms_stage2 = sum(raw2[1:]) / len(raw2[1:])  # steady state, like `ms` above
    per2 = sorted(raw2[1:])
    q2_ = lambda f: per2[min(len(per2) - 1, int(f * len(per2)))]  # noqa: E731

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
        o_ = torch_trunk()
        b.record()
        o_.sum().backward()
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
    flop_alg = 2.0 * n * (96 * 256 + 6 * 256 * 256 + 352 * 256 + 256 * 32)  # one fp32-accurate product per weight
    ach_issued = 3 * flop_alg / (ms_deform_fwd * 1e-3) / 1e12
    ach_alg = flop_alg / (ms_deform_fwd * 1e-3) / 1e12
    return {
        "ms": ms_full, "ms_quantiles": q_full, "ms_without_deform": ms_plain, "ms_without_deform_quantiles": q_plain,
        "deform_fwd_ms": ms_deform_fwd, "deform_bwd_ms": ms_deform_bwd,
        "stage2_ms": ms_stage2, "stage2_quantiles": {"p10": q2_(0.1), "p50": q2_(0.5), "p90": q2_(0.9)},
        "stage2_controlled_gaussians": int(part.numel()),
        "gaussians": n, "visible": n_vis, "launches_per_iter": launches, "steps": steps, "warmup": warmup,
        "host_enqueue_ms_per_iter": host_ms, "diagnostics": diag, "network_backward_paths": dict(__import__("freegaussian_b200.deform", fromlist=["STATS"]).STATS),
        "config": "stage-1 step: DeformNetwork(is_blender=True) -> RGB+ED render -> blend+L1+SSIM -> backward -> Adam; "
                  "stage2_ms: ControlNetwork on the controlled subset -> the same render / loss / backward / Adam",
        "library_kernels": "torch.optim.Adam(fused=True) for the 0.6 M network weights, torch ops for the one-row time branch "
                           "and for sigmoid / exp / index_copy glue; everything else is this repo's kernels",
        "deform_roofline": {"bound": "tensor", "achieved": ach_issued, "peak": tf32_peak, "unit": "TFLOP/s",
                            "frac": ach_issued / tf32_peak, "achieved_algorithmic": ach_alg,
                            "frac_algorithmic": ach_alg / tf32_peak,
                            "note": "forward.  `achieved` counts the tf32 MMA flops ISSUED (3xTF32: three products per fp32-accurate "
                                    "product); `achieved_algorithmic` counts each product once; peak = measured bf16 GEMM peak / 2"},
        "torch_fp32_network": {"fwd_ms": torch_fwd, "bwd_ms": torch_bwd,
                               "note": "the same network with torch.nn.functional.linear in fp32 on this GPU (what the reference runs)"},
    }


# ------------------------------------------------------------------------------ CPU arm
def cpu_arm(args, steps: int, warmup: int, scene=None):
    """The reference algorithm (oracle restatement of the gsplat path -- gsplat itself cannot be installed here,
    SURVEY.md 8(c)) on the host cores, as a bounded sample of the SAME step: same scene and camera, ALL Gaussians projected,
    SH-evaluated and back-propagated (the per-Gaussian part, timed in full: t_gauss), and a CPU_CROP x CPU_CROP centre crop
    of the frame tiled, sorted and composited fwd+bwd (the per-pixel part: t_pix).  A full frame costs the per-Gaussian
    part once and the per-pixel part W*H / crop-area times, so the full-frame throughput is extrapolated as
        value = W*H / (t_gauss + t_pix * W*H / (cw*ch))
    (the centre crop is denser than the frame's average, so this favours neither arm by much; `crop_value` = cw*ch / step
    time is what round 1 reported and understates the CPU by amortising t_gauss over the crop only).  Nothing of the
    product library is loaded by this arm: the scene's 3-NN scales come from the reference's own sklearn call."""
    from oracle import knn as oknn
    from oracle import render as O
    from freegaussian_b200.scenes import make_scene

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, W, H = WORKLOADS[args.workload]
    if scene is None:
        knn3 = lambda m: torch.from_numpy(oknn.reference_knn(m.numpy(), 3)[0])  # noqa: E731
        scene = make_scene(n, W, H, n_views=N_VIEW_POOL, recipe=args.recipe, seed=0, knn3=knn3)
    sc = scene
    cw, ch = min(CPU_CROP, W), min(CPU_CROP, H)
    K = sc.Ks[:1].clone()
    K[:, 0, 2] -= (W - cw) / 2
    K[:, 1, 2] -= (H - ch) / 2
    vm = sc.viewmats[:1]
    rgb, depth, flow = loss_weights(ch, cw, 1)
    w_rgbd, w_flow = decode_weights(rgb, depth, flow)
    names = ("means", "quats", "scales", "opacities", "sh", "means_next")
    params = [getattr(sc, k).detach().clone().requires_grad_(True) for k in names]
    means, quats, scales, opac, sh, mnext = params
    t_total, t_gauss = [], []
    n_isect = 0
    for i in range(warmup + steps):
        for p in params:
            p.grad = None
        t0 = time.perf_counter()
        r, a, m = O.rasterization(means, quats, scales, opac, sh, vm, K, cw, ch, near_plane=0.01, far_plane=1e10,
                                  render_mode="RGB+ED", sh_degree=3, means_next=mnext)
        loss = (r * w_rgbd).sum() + (m["flow"] * w_flow).sum()
        loss.backward()
        t1 = time.perf_counter()
        # the per-Gaussian part alone: projection, SH colours and the frame t+1 projection, forward + backward
        for p in params:
            p.grad = None
        radii, means2d, depths, conics, _, _ = O.fully_fused_projection(means, quats, scales, vm, K, cw, ch, 0.3, 0.01, 1e10, 0.0)
        dirs = means[None] - torch.inverse(vm)[:, :3, 3][:, None]
        cols = torch.clamp_min(O.spherical_harmonics(3, dirs, sh[None], masks=radii > 0) + 0.5, 0.0)
        uv_next, _ = O.project_points(mnext, vm, K)
        (means2d.sum() + depths.sum() + conics.sum() + cols.sum() + (uv_next - means2d).sum()).backward()
        t2 = time.perf_counter()
        n_isect = m["flatten_ids"].numel()
        if i >= warmup:
            t_total.append(t1 - t0)
            t_gauss.append(t2 - t1)
    tt, tg = sum(t_total) / len(t_total), sum(t_gauss) / len(t_gauss)
    tg = min(tg, tt)
    tp = tt - tg
    area = W * H / (cw * ch)
    full_s = tg + tp * area
    val = W * H / full_s / 1e6
    sample = (f"{steps} step(s) of fwd+bwd: all {n} Gaussians projected (per-Gaussian part {tg:.2f} s), a {cw}x{ch} centre crop of "
              f"the {W}x{H} frame composited ({n_isect} intersections, per-pixel part {tp:.2f} s); full frame extrapolated as "
              f"t_gauss + t_pix x {area:.1f} = {full_s:.1f} s/step (crop-extrapolated)")
    base = {"value": val, "unit": "MPix/s", "cores": cores, "kind": "port", "sample": sample,
            "crop_value": cw * ch / tt / 1e6, "t_gauss_s": tg, "t_pix_s": tp, "extrapolated_full_frame_s": full_s}
    return {"cpu_baseline": base, "extrapolated_ms_per_step": full_s * 1e3, "sample_ms_per_step": (tt + tg) * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = max(args.warmup, 3)
    res = cpu_arm(args, steps=args.steps, warmup=warm)
    val = res["cpu_baseline"]["value"]
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": val, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        # ms_per_step: wall time of one bounded-sample step as executed; value: the full-frame throughput extrapolated from it
        "ms_per_step": res["sample_ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, args.recipe, args.views_per_gpu),
                   "note": "CPU arm: each step is a bounded sample of the workload (per-Gaussian part in full, per-pixel part on a "
                           "centre crop), value = crop-extrapolated full-frame throughput; see cpu_baseline.sample",
                   "extrapolated_full_frame_ms_per_step": res["extrapolated_ms_per_step"]},
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
