#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "chain or tma or tile or evaluate" > $O/r02n_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest_new.log
tail -3 $O/r02n_pytest_new.log
timeout 200 python tools/probe_timeline.py > $O/r02n_timeline.log 2>&1; grep "^item" $O/r02n_timeline.log | sed -n '2p;12p;13p' | cut -c1-330
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02n_bench.json 2> $O/r02n_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02n_bench.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["hbm_regime"]["tile_W"], d["hbm_regime"]["h_apply"]["us"])
PY
