#!/usr/bin/env python
"""fg_xchg_allreduce_f32 alone (no kernel beside it) against NCCL's all-reduce, same buffer sizes, under torchrun:
   python -m torch.distributed.run --nproc-per-node G tools/bench_xchg_allreduce.py
Sweeps the CTA count of the in-switch path.  Prints ms, algorithm and bus bandwidth (rank 0)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freegaussian_b200 import _lib  # noqa: E402
from freegaussian_b200.dist import ViewShardedExchange  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
L = _lib.lib()
xc = ViewShardedExchange()


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for floats in (17_000_000, 62_000_000):  # geometry gradients of 1 M Gaussians / the dense arena of round 1
    mb = floats * 4 / 1e6
    x = torch.ones(floats, device=dev)
    ms = timed(lambda: dist.all_reduce(x))
    if rank == 0:
        print(f"NCCL all_reduce        {mb:6.0f} MB world {world}: {ms:.3f} ms  bus {mb / ms * 2 * (world - 1) / world:.0f} GB/s", flush=True)
    arena = xc.arena(floats, dev)
    arena.fill_(1.0)
    st = torch.cuda.current_stream().cuda_stream

    def ours():
        xc.epoch += 1
        _lib.check(L.fg_xchg_allreduce_f32(xc._peers, xc.arena_off, floats // 4 * 4, xc.epoch, 1, st))

    for blocks in ((16, 32, 64, 128) if xc.multicast else (0,)):
        if blocks:
            _lib.check(L.fg_set_option(b"xchg_ar_blocks", blocks))
        arena.fill_(1.0)
        ms = timed(ours)
        if rank == 0:
            path = f"multimem {blocks:3d} CTAs" if xc.multicast else "peer ld/st       "
            print(f"fg_xchg {path} {mb:6.0f} MB world {world}: {ms:.3f} ms  bus {mb / ms * 2 * (world - 1) / world:.0f} GB/s", flush=True)
dist.barrier()
dist.destroy_process_group()
