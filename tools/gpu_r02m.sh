#!/bin/bash
O=gpurun_out; mkdir -p $O
for u in 1 2; do
  FH_PROBE_DEFS="-DPAIR_UNROLL=$u" timeout 200 python tools/probe_timeline.py > $O/r02m_timeline_u$u.log 2>&1
  echo "== unroll $u"; grep "^item" $O/r02m_timeline_u$u.log | sed -n '2p;12p' | cut -c1-330
done
