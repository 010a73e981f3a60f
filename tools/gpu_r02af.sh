#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q > $O/r02af_pytest.log 2>&1; echo "rc=$?" >> $O/r02af_pytest.log
tail -4 $O/r02af_pytest.log
timeout 600 python tools/sweep_roofline.py --lattices 3x4 2>&1 | tail -6
timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-hbm-regime 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d['h_eval_ms'])"
