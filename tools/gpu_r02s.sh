#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tile_passes or apply_table_22" > $O/r02s_pytest_k2.log 2>&1; echo "rc=$?" >> $O/r02s_pytest_k2.log
tail -3 $O/r02s_pytest_k2.log
for lat in 3x4 2x6; do
  timeout 200 python tools/run_k2.py $lat 10 2>&1 | tail -1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_table_pass' -c 2 -o $O/r02s_k2 -f python tools/run_k2.py 3x4 1 > $O/r02s_ncu.log 2>&1
ncu -i $O/r02s_k2.ncu-rep --page raw --csv > $O/r02s_k2_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02s_k2_raw.csv
ncu -i $O/r02s_k2.ncu-rep --page details 2>/dev/null | grep -E "Issue Slots Busy|No Eligible|Executed Ipc Active|Warp Cycles Per Issued|This stall type|cycles being stalled" | head -12
