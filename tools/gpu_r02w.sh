#!/bin/bash
O=gpurun_out; mkdir -p $O
for tb in 10 12 13; do
FHSIM_TILE_BITS=$tb timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-hbm-regime 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tile_bits=$tb', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d['roofline']['kernel_ms'])"
done
