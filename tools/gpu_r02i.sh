#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "chain or tma or tile or evaluate" > $O/r02i_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest_new.log
tail -4 $O/r02i_pytest_new.log
if grep -q "pytest rc=0" $O/r02i_pytest_new.log; then
timeout 300 python tools/probe_timeline.py --rebuild > $O/r02i_timeline_tma.log 2>&1
grep "^item" $O/r02i_timeline_tma.log | cut -c1-400
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02i_bench.json 2> $O/r02i_bench.err; echo "rc=$?" >> $O/r02i_bench.err
FHSIM_CHAIN=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-hbm-regime > $O/r02i_bench_chain.json 2> $O/r02i_bench_chain.err
cut -c1-330 $O/r02i_bench.json; echo; cut -c1-330 $O/r02i_bench_chain.json; echo; tail -3 $O/r02i_bench.err
fi
