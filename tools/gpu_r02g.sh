#!/bin/bash
# r02 session G: 4-warp CTAs with 4 independent pairs per thread
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "chain or tma or tile or evaluate or dressing" > $O/r02g_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest_new.log
tail -8 $O/r02g_pytest_new.log
if grep -q "pytest rc=0" $O/r02g_pytest_new.log; then
timeout 300 python tools/probe_timeline.py --rebuild > $O/r02g_timeline_tma.log 2>&1; echo "rc=$?" >> $O/r02g_timeline_tma.log
grep "^item\|rc=" $O/r02g_timeline_tma.log | cut -c1-420
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02g_bench_chain.json 2> $O/r02g_bench_chain.err; echo "rc=$?" >> $O/r02g_bench_chain.err
FHSIM_NO_CHAIN=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-hbm-regime > $O/r02g_bench_nochain.json 2> $O/r02g_bench_nochain.err
cut -c1-330 $O/r02g_bench_chain.json; echo; cut -c1-330 $O/r02g_bench_nochain.json; echo; tail -3 $O/r02g_bench_chain.err
fi
