#!/bin/bash
# r02 session B: first run of the TMA tile kernel: targeted tests, then the suite, bench TMA vs LDG, probe timings
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tile" > $O/r02b_pytest_tile.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest_tile.log
tail -15 $O/r02b_pytest_tile.log
if grep -q "pytest rc=0" $O/r02b_pytest_tile.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
  tail -5 $O/r02b_pytest.log
  timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02b_bench_tma.json 2> $O/r02b_bench_tma.err; echo "rc=$?" >> $O/r02b_bench_tma.err
  FHSIM_TILE_LDG=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $O/r02b_bench_ldg.json 2> $O/r02b_bench_ldg.err
  cut -c1-400 $O/r02b_bench_tma.json; echo; cut -c1-400 $O/r02b_bench_ldg.json; echo
  tail -3 $O/r02b_bench_tma.err
fi
timeout 120 tools/probes/probe_tma > $O/r02b_probe_tma.log 2>&1; echo "probe rc=$?" >> $O/r02b_probe_tma.log
cat $O/r02b_probe_tma.log
