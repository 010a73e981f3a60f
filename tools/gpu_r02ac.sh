#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02final_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02final_pytest.log
tail -8 $O/r02final_pytest.log; grep -E "^FAILED|^ERROR" $O/r02final_pytest.log | head
