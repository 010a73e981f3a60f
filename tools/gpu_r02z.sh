#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02z_pytest_all.log 2>&1; echo "rc=$?" >> $O/r02z_pytest_all.log
tail -6 $O/r02z_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
