#!/bin/bash
# quick single-GPU bench line (no CPU legs): python bench.py --no-cpu-baseline -> gpurun_out/quick_bench.json
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > $O/quick_bench.json 2> $O/quick_bench.err; echo "bench rc=$?"; tail -2 $O/quick_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], "h_eval", d["h_eval_ms"], "train", d["train_step"])
PY
