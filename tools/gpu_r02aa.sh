#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q > $O/r02aa_pytest.log 2>&1; echo "rc=$?" >> $O/r02aa_pytest.log
tail -12 $O/r02aa_pytest.log
timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-hbm-regime > $O/r02aa_bench.json 2> $O/r02aa_bench.err; echo "bench rc=$?"; tail -3 $O/r02aa_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02aa_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"], d["config"]["path"])
print("h_evals", d["h_evals_per_s"], d["h_eval_ms"], d["h_eval_launches"])
PY
