#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/r02final_bench.json 2> $O/r02final_bench.err; echo "bench rc=$?"; tail -2 $O/r02final_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02final_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launches_per_step"])
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","kernel_ms","launches_per_step","share_of_step","traffic")})
PY
