#!/bin/bash
# Run on the GPU box (gpurun):  bash tools/capture_profiles.sh <tag>
# Writes gpurun_out/<tag>_launches.csv (per-launch durations of a short bench run) and one
# `ncu --set full` report per hot kernel; summarise here with tools/ncu_summary.py into profiles/.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-train-iter"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH > $OUT/${TAG}_launches_bench.log 2>&1
for k in rasterize_bwd2_kernel rasterize_fwd_kernel project_bwd_kernel sh_bwd_kernel project_fwd_kernel fine_bin_kernel ranked_emit_kernel bin_count_cells_kernel; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 --launch-count 1 \
        -f -o $OUT/${TAG}_$k $BENCH > $OUT/${TAG}_$k.log 2>&1
done
# k-NN at the cfg5 shape (3 M points, k = 16): the query kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_query_kernel --launch-skip 1 --launch-count 1 \
    -f -o $OUT/${TAG}_knn_query_kernel python tools/knn_once.py > $OUT/${TAG}_knn_query_kernel.log 2>&1
# the exchange kernels in a one-rank group (FG_XCHG_SOLO): 8 views of cfg3 published and summed on one GPU -- counters of the
# kernels themselves (instruction mix, DRAM traffic); the NVLink side is measured by bench.py at N > 1, never under ncu
for k in sh_bwd_views_kernel allreduce_kernel; do
    FG_XCHG_SOLO=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 --launch-count 1 \
        -f -o $OUT/${TAG}_$k python tools/xchg_solo.py > $OUT/${TAG}_$k.log 2>&1
done
ls -la $OUT | tail -20
