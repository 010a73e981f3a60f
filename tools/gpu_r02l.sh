#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "window or apply_table or fused_runs" > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest.log
tail -4 $O/r02l_pytest.log
for sb in 0 2 3 4 5; do
  FHSIM_K2_SLOW_BITS=$sb timeout 200 python tools/run_k2.py 3x4 > $O/r02l_k2_sb$sb.log 2>&1; echo "slow_bits=$sb: $(tail -1 $O/r02l_k2_sb$sb.log | cut -c1-200)"
done
python tools/profile_24q.py 3x4 > $O/r02l_plain24.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_apply_table --csv --log-file $O/r02l_k2_ncu.csv python tools/profile_24q.py 3x4 > $O/r02l_ncu.log 2>&1
grep -v "^==" $O/r02l_k2_ncu.csv | tail -8 | cut -c1-250
