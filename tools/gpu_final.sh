#!/bin/bash
# Slim end-of-session GPU validation: full parity suite, bench line, ncu launch list.  Outputs -> gpurun_out/<tag>_*
TAG=${1:-r01f}
O=gpurun_out
mkdir -p $O
date +%T
timeout 200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -2 $O/${TAG}_pytest.log; date +%T
timeout 170 python bench.py --steps 1000 --warmup 20 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
date +%T
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/${TAG}_ncu_launches.log 2>&1
date +%T
cut -c1-300 $O/${TAG}_bench.json
