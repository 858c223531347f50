#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/r1_launches.csv            # per-kernel time shares
  python tools/ncu_summary.py raw gpurun_out/r1_rasterize_bwd_kernel.ncu-rep # key counters of a --set full capture
"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    # tensor-core kernels (csrc/mlp.cu)
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[hdr]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e6:.3f} ms total (ncu-serialised, cold cache: compare SHARES)")
    print(f"{'total ms':>10} {'count':>6} {'avg us':>10} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{v[1]/1e6:10.3f} {v[0]:6d} {v[1]/v[0]/1e3:10.1f} {100*v[1]/tot:6.1f}%  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    for r in rows[2:]:
        print(f"# {path}: {r[h.index('Kernel Name')][:90]}  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
        for w in WANT:
            if w in h:
                print(f"  {w:75s} {rows[1][h.index(w)]:>16s} {r[h.index(w)]}")
        stalls = [(float(r[i]), c) for i, c in enumerate(h) if c.startswith("smsp__average_warp") and "issue_stalled" in c and c.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        if not stalls:
            stalls = [(float(r[i].replace(",", "")), c) for i, c in enumerate(h) if "issue_stalled" in c and c.endswith(".pct") and r[i] not in ("", "n/a")]
        for v, c in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {c:69s} {v:12.3f}")


def source(path, top=25):
    """Hottest CUDA source lines by warp-stall samples (needs -lineinfo and --import-source on)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    data, h, fname = [], None, ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            h = r
            ci = h.index("Warp Stall Sampling (All Samples)")
            ii = h.index("Instructions Executed")
        elif h and r and r[0].isdigit() and len(r) > ci:
            try:
                data.append((float(r[ci] or 0), float(r[ii] or 0), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot = sum(d[0] for d in data) or 1
    print(f"# {path}: hottest source lines (share of warp-stall samples | warp instructions executed)")
    for smp, ins, fn, ln, src in sorted(data, reverse=True)[:top]:
        print(f"{100*smp/tot:6.2f}% {ins:13.0f}  {fn}:{ln:<4d} {src[:105]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "source": source}[sys.argv[1]](sys.argv[2])
