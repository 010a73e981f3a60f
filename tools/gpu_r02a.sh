#!/bin/bash
# r02 session A: TMA probe, parity tests (new value-level cfg2/3/4 tests), baseline bench before kernel work
O=gpurun_out; mkdir -p $O
timeout 120 tools/probes/probe_tma > $O/r02a_probe_tma.log 2>&1; echo "probe rc=$?" >> $O/r02a_probe_tma.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02a_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench rc=$?" >> $O/r02a_bench.err
cat $O/r02a_probe_tma.log; tail -5 $O/r02a_pytest.log; cut -c1-600 $O/r02a_bench.json
