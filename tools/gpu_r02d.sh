#!/bin/bash
# r02 session D: table-driven op loop + bulk-copied descriptors; chain vs per-run launches; timelines
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "chain or sector or tma or tile" > $O/r02d_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest_new.log
tail -25 $O/r02d_pytest_new.log
if grep -q "pytest rc=0" $O/r02d_pytest_new.log; then
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02d_bench_chain.json 2> $O/r02d_bench_chain.err; echo "rc=$?" >> $O/r02d_bench_chain.err
FHSIM_NO_CHAIN=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-hbm-regime > $O/r02d_bench_nochain.json 2> $O/r02d_bench_nochain.err
cut -c1-330 $O/r02d_bench_chain.json; echo; cut -c1-330 $O/r02d_bench_nochain.json; echo; tail -3 $O/r02d_bench_chain.err
timeout 300 python tools/probe_timeline.py --rebuild > $O/r02d_timeline_tma.log 2>&1; echo "rc=$?" >> $O/r02d_timeline_tma.log
grep -v warning $O/r02d_timeline_tma.log | cut -c1-420 | tail -16
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
tail -5 $O/r02d_pytest.log
fi
