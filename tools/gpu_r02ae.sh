#!/bin/bash
O=gpurun_out; mkdir -p $O
FHSIM_SECTOR_PREFIX=1 timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q -k "dense or screening_vs_oracle" > $O/r02ae_pytest.log 2>&1; echo "rc=$?" >> $O/r02ae_pytest.log
tail -4 $O/r02ae_pytest.log
for pre in 0 1; do
  if [ $pre = 1 ]; then export FHSIM_SECTOR_PREFIX=1; else unset FHSIM_SECTOR_PREFIX; fi
  timeout 600 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-hbm-regime 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('prefix=$pre', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d['h_eval_ms'], d['h_eval_launches'])"
done
