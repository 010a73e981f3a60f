#!/bin/bash
# the whole -m gpu suite + smoke -> gpurun_out/tests_all_*.log
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/tests_all_pytest.log 2>&1; echo "pytest rc=$?" >> $O/tests_all_pytest.log
tail -4 $O/tests_all_pytest.log; grep -E "^FAILED|^ERROR" $O/tests_all_pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
