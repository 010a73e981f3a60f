#!/bin/bash
# sector-path parity tests + a quick bench line
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q > $O/quick_sector_pytest.log 2>&1; echo "rc=$?" >> $O/quick_sector_pytest.log
tail -3 $O/quick_sector_pytest.log
bash tools/gpu_quick_bench.sh
