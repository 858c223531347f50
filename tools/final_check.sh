#!/bin/bash
# Run on the GPU box (gpurun):  bash tools/final_check.sh <tag>
# Everything the round-end driver runs (GPU tests, smoke, both bench arms) plus the ncu launch lists of the bench
# command and of one deformation-network step; outputs in gpurun_out/<tag>_*.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
(time timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > $OUT/${TAG}_pytest.log 2>&1
tail -6 $OUT/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
tail -2 $OUT/${TAG}_smoke.log
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -2 $OUT/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-iter > $OUT/${TAG}_launches_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches_deform.csv \
    python tools/prof_deform.py > $OUT/${TAG}_prof_deform.log 2>&1
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"])
t = d["train_iter"]
print({k: round(t[k], 3) for k in ("ms", "ms_without_deform", "deform_fwd_ms", "deform_bwd_ms", "stage2_ms")},
      t["stage2_controlled_gaussians"], t["torch_fp32_network"])
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
cut -c1-300 $OUT/${TAG}_bench_reference.json
