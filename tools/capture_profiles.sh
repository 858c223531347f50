#!/bin/bash
# Run on the GPU box (gpurun):  bash tools/capture_profiles.sh <tag>
# Writes gpurun_out/<tag>_launches.csv (per-launch durations of a short bench run) and one
# `ncu --set full` report per hot kernel; summarise here with tools/ncu_summary.py into profiles/.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH > $OUT/${TAG}_launches_bench.log 2>&1
for k in rasterize_bwd2_kernel rasterize_fwd_kernel project_bwd_kernel sh_bwd_kernel project_fwd_kernel fine_bin_kernel; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 --launch-count 1 \
        -f -o $OUT/${TAG}_$k $BENCH > $OUT/${TAG}_$k.log 2>&1
done
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
ls -la $OUT | tail -20
