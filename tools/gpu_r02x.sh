#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py tests/test_gpu_sharded.py -m gpu -x -q > $O/r02x_pytest.log 2>&1; echo "rc=$?" >> $O/r02x_pytest.log
tail -6 $O/r02x_pytest.log
