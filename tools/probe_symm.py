"""Probe (torchrun, >= 2 ranks): is torch symmetric memory usable on this box (peer pointers, NVLS multicast
pointer), and what do the NCCL collectives the exchange step could use cost at the sizes in question?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm

def p(*a):
    if rank == 0:
        print(*a, flush=True)

try:
    t = symm.empty(64 << 20, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    p("symm ok: world", h.world_size, "buffer_ptrs", [hex(x) for x in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr or 0),
      "has_multicast", symm._SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, lr) if hasattr(symm._SymmetricMemory, "has_multicast_support") else None,
      "signal_pad", [hex(x) for x in h.signal_pad_ptrs], "pad size", h.signal_pad_size, "backend", symm.get_backend(dev))
except Exception as e:
    p("symm FAILED:", type(e).__name__, e)

def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for mb in (1, 4, 14, 56, 236):
    x = torch.zeros(mb * 250_000, device=dev)
    ms = timeit(lambda: dist.all_reduce(x))
    p(f"nccl all_reduce {mb} MB: {ms:.3f} ms  algbw {mb/ms:.0f} GB/s")
for mb in (1, 4, 12):
    x = torch.zeros(mb * 250_000, device=dev)
    out = torch.empty(world * x.numel(), device=dev)
    ms = timeit(lambda: dist.all_gather_into_tensor(out, x))
    p(f"nccl all_gather {mb} MB/rank: {ms:.3f} ms")
dist.barrier()
dist.destroy_process_group()
