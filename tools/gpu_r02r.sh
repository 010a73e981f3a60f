#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sector.py -m gpu -x -q > $O/r02r_pytest_sector.log 2>&1; echo "rc=$?" >> $O/r02r_pytest_sector.log
tail -15 $O/r02r_pytest_sector.log
