#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q > $O/r02u_pytest_sector.log 2>&1; echo "rc=$?" >> $O/r02u_pytest_sector.log
tail -4 $O/r02u_pytest_sector.log
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02u_bench.json 2> $O/r02u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02u_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms"], d["roofline_k3"]["kernel_ms"], d["hbm_regime"]["h_apply"]["us"])
PY
FHSIM_NO_SECTOR_POOL=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no sector pool:', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
