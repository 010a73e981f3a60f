#!/bin/bash
for c in 0 1; do
  if [ $c = 1 ]; then export FHSIM_CHAIN=1; else unset FHSIM_CHAIN; fi
  timeout 300 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-hbm-regime 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chain=$c', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d['train_step']['ms'] if d.get('train_step') else None)"
done
