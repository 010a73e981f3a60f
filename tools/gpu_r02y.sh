#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q > $O/r02y_pytest.log 2>&1; echo "rc=$?" >> $O/r02y_pytest.log
tail -6 $O/r02y_pytest.log
timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline > $O/r02y_bench.json 2> $O/r02y_bench.err; echo "bench rc=$?"; tail -3 $O/r02y_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02y_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d["config"]["path"])
print("h_evals", d["h_evals_per_s"], d["h_eval_ms"], d["h_eval_launches"])
for k,v in d["hbm_regime"].items():
    if isinstance(v, dict): print(k, v["us"], v["frac"])
PY
