#!/bin/bash
# r02 session C: chain kernel + sector Lanczos first runs, timelines of the TMA tile kernel, bench
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "chain or sector or tma" > $O/r02c_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest_new.log
tail -25 $O/r02c_pytest_new.log
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-hbm-regime > $O/r02c_bench_chain.json 2> $O/r02c_bench_chain.err; echo "rc=$?" >> $O/r02c_bench_chain.err
FHSIM_NO_CHAIN=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-hbm-regime > $O/r02c_bench_nochain.json 2> $O/r02c_bench_nochain.err
cut -c1-330 $O/r02c_bench_chain.json; echo; cut -c1-330 $O/r02c_bench_nochain.json; echo; tail -3 $O/r02c_bench_chain.err
timeout 300 python tools/probe_timeline.py --rebuild > $O/r02c_timeline_tma.log 2>&1; echo "rc=$?" >> $O/r02c_timeline_tma.log
cat $O/r02c_timeline_tma.log | cut -c1-400
timeout 300 python tests/perf_lanczos.py --json $O/r02c_lanczos.json > $O/r02c_lanczos.log 2>&1; tail -8 $O/r02c_lanczos.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -5 $O/r02c_pytest.log
timeout 600 python tools/lanczos_4x4.py --json $O/r02c_lanczos_4x4.json > $O/r02c_lanczos_4x4.log 2>&1; echo "rc=$?" >> $O/r02c_lanczos_4x4.log; tail -3 $O/r02c_lanczos_4x4.log | cut -c1-600
