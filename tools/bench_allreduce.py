#!/usr/bin/env python
"""all-reduce of the 1 M-Gaussian gradient arena (59 floats per Gaussian + tail) under torchrun; prints ms and GB/s."""
import os, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.ones(62_000_000, device="cuda")
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
if rank == 0:
    print(f"{os.environ.get('TAG','default')}: world {world}: {ms:.3f} ms, algbw {x.numel()*4/ms/1e6:.0f} GB/s, busbw {x.numel()*4/ms/1e6*2*(world-1)/world:.0f} GB/s")
dist.destroy_process_group()
