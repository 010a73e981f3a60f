#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > $O/r02v_bench.json 2> $O/r02v_bench.err; echo "bench rc=$?"; tail -3 $O/r02v_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02v_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["launches_per_step"], d["config"]["path"])
print("k3", d["roofline_k3"]["kernel_ms"], d["roofline_k3_full_space"]["kernel_ms"], "tile", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"])
for k,v in d["hbm_regime"].items():
    if isinstance(v, dict): print(k, v["us"], v["frac"])
print("h_evals", d["h_evals_per_s"], d["h_eval_ms"])
PY
