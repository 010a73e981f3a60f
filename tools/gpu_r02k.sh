#!/bin/bash
O=gpurun_out; mkdir -p $O
FHSIM_NO_RUNS=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02k_bench_noruns.json 2> $O/r02k_bench_noruns.err
python - <<'PY'
import json
for f in ("gpurun_out/r02k_bench_noruns.json",):
    d=json.load(open(f)); print(f, d["ms_per_step"], d["hbm_regime"]["tile_W"])
PY
