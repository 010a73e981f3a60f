#!/bin/bash
# Round-2 final evidence session: parity tests, smoke, bench (both arms), ncu launch list + full captures (18 q step, 24 q
# kernels), roofline sweep incl. 3x5 (30 qubits).  Outputs -> gpurun_out/<tag>_*; summaries are copied into profiles/.
TAG=${1:-r02final}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" >> $O/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
# launch list (per-launch durations, cold cache, serialised)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/${TAG}_ncu_launches.log 2>&1
# full capture of the kernels of the 18-qubit step (skip build / warm-up launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_tma|k_sector_gemm|k_sector_happly|k_sector_pool' -s 60 -c 40 \
    -o $O/${TAG}_full18 -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/${TAG}_ncu_full18.log 2>&1
ncu -i $O/${TAG}_full18.ncu-rep --page raw --csv > $O/${TAG}_full18_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_full18_raw.csv > $O/${TAG}_ncu_full_18q.md
# the kernels at 24 qubits (state 256 MiB > L2): dram bytes per launch for the HBM rooflines
timeout 900 ncu --set full --clock-control none -k regex:'k_pool|k_pair|k_apply_table|k_table_pass|k_diag_tab|k_tile|k_sector' \
    -o $O/${TAG}_full24 -f python tools/profile_24q.py 3x4 > $O/${TAG}_ncu_full24.log 2>&1
ncu -i $O/${TAG}_full24.ncu-rep --page raw --csv > $O/${TAG}_full24_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_full24_raw.csv > $O/${TAG}_ncu_full_24q.md
rm -f $O/${TAG}_full24.ncu-rep $O/${TAG}_full18.ncu-rep $O/${TAG}_full24_raw.csv $O/${TAG}_full18_raw.csv
timeout 900 python tools/sweep_roofline.py --lattices 3x3,3x4,2x7,3x5 --json $O/${TAG}_sweep.json > $O/${TAG}_sweep.log 2>&1
tail -3 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log; cut -c1-600 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err; cut -c1-400 $O/${TAG}_bench_ref.json; tail -12 $O/${TAG}_sweep.log
