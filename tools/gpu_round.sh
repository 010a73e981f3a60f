#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list + full captures. Outputs -> gpurun_out/<tag>_*
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
set -x; PS4="+ $(date +%T) "
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" >> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
# launch list (per-launch durations, cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/${TAG}_ncu_launches.log 2>&1
# full capture of the three dominant kernels of the 18-qubit step (skip the first 60 launches = build/warm-up)
ncu --set full --clock-control none --import-source on -k regex:'k_pool$|k_tile|k_apply_table' -s 60 -c 24 \
    -o $O/${TAG}_full18 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/${TAG}_ncu_full18.log 2>&1
ncu -i $O/${TAG}_full18.ncu-rep --page raw --csv > $O/${TAG}_full18_raw.csv 2>/dev/null
# the same kernels at 24 qubits (state 256 MiB > L2): dram bytes per launch for the HBM rooflines
ncu --set full --clock-control none -k regex:'k_pool|k_pair|k_apply_table|k_diag_tab|k_tile' \
    -o $O/${TAG}_full24 -f python tools/profile_24q.py 3x4 > $O/${TAG}_ncu_full24.log 2>&1
ncu -i $O/${TAG}_full24.ncu-rep --page raw --csv > $O/${TAG}_full24_raw.csv 2>/dev/null
rm -f $O/${TAG}_full24.ncu-rep $O/${TAG}_full18.ncu-rep      # keep gpurun_out small: the raw CSVs are what we read
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json
