#!/bin/bash
O=gpurun_out; mkdir -p $O
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/r02j_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_tma -s 56 -c 4 -o $O/r02j_tileW -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/r02j_ncu.log 2>&1
echo "ncu rc=$?"; ls -la $O/r02j_tileW.ncu-rep
