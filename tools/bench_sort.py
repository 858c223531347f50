#!/usr/bin/env python
"""Micro-benchmark of the hand-written radix sort (GPU): GB/s per configuration.
  python tools/bench_sort.py [n]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freegaussian_b200 import _build, _lib
_build.build()
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 33_000_000
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(L.fg_radix_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")


def run(name, keys, end_bit, u64):
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    kb, vb = torch.empty_like(keys), torch.empty_like(vals)
    fn = L.fg_radix_sort_pairs_u64_u32 if u64 else L.fg_radix_sort_pairs_u32_u32
    sel = ctypes.c_int(0)
    ts = []
    for it in range(6):
        ka, va = keys.clone(), vals.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(n, ka.data_ptr(), va.data_ptr(), kb.data_ptr(), vb.data_ptr(), end_bit, ws.data_ptr(), ws.numel(), ctypes.byref(sel), st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    passes = (end_bit + 7) // 8
    ksz = 8 if u64 else 4
    byts = n * (ksz + passes * 2 * (ksz + 4))
    print(f"{name:34s} n={n} passes={passes} {t:8.3f} ms  {byts/t/1e6:8.1f} GB/s  ({t/passes*1e3:7.1f} us/pass incl. hist)")


g = torch.Generator(device="cuda").manual_seed(0)
run("u32 random 13 bits", torch.randint(0, 8160, (n,), generator=g, device="cuda", dtype=torch.int32), 13, False)
run("u32 random 32 bits", torch.randint(-2**31, 2**31 - 1, (n,), generator=g, device="cuda", dtype=torch.int32), 32, False)
# emission-like: runs of consecutive tile ids (rows of a splat's rectangle)
base = torch.randint(0, 8160 - 12, (n // 11 + 1,), generator=g, device="cuda", dtype=torch.int32)
em = (base[:, None] + torch.arange(11, device="cuda", dtype=torch.int32)[None]).reshape(-1)[:n].contiguous()
run("u32 emission-like runs 13 bits", em, 13, False)
run("u64 random 46 bits", torch.randint(0, 2**46, (n,), generator=g, device="cuda", dtype=torch.int64), 46, True)
