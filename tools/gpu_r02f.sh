#!/bin/bash
# ncu --set full with source on the TMA tile kernel (18-qubit bench step, one launch per run)
O=gpurun_out; mkdir -p $O
export FHSIM_NO_CHAIN=1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/r02f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_tma -s 45 -c 15 -o $O/r02f_tile18 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/r02f_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02f_ncu.log; ls -la $O/r02f_tile18.ncu-rep
