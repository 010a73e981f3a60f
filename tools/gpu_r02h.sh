#!/bin/bash
O=gpurun_out; mkdir -p $O
for v in 0 1 2 3 4; do
  FH_OPLOOP_VARIANT=$v timeout 200 python tools/probe_timeline.py > $O/r02h_timeline_v$v.log 2>&1
  echo "== variant $v"; grep "^item" $O/r02h_timeline_v$v.log | sed -n '2p;12p' | cut -c1-330
done
