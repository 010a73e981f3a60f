#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_table_pass|k_apply_table4' -c 4 -o $O/r02t_k2 -f python tools/run_k2.py 3x4 1 > $O/r02t_ncu.log 2>&1
ncu -i $O/r02t_k2.ncu-rep --page raw --csv > $O/r02t_k2_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02t_k2_raw.csv
ncu -i $O/r02t_k2.ncu-rep --page details 2>/dev/null | grep -E "Issue Slots Busy|No Eligible|Eligible Warps|Executed Ipc|L1/TEX Hit|Shared.*Bank|Warp Cycles Per Issued|Stall|FP64|Registers Per|Theoretical Occ|Achieved Occ|Local|spill" | head -60
