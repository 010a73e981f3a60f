#!/bin/bash
# r02 session C: in-kernel timelines of the TMA tile kernel vs the register-staged one, per-item launch times
O=gpurun_out; mkdir -p $O
timeout 300 python tools/probe_timeline.py > $O/r02c_timeline_tma.log 2>&1; echo "rc=$?" >> $O/r02c_timeline_tma.log
FHSIM_TILE_LDG=1 timeout 300 python tools/probe_timeline.py > $O/r02c_timeline_ldg.log 2>&1; echo "rc=$?" >> $O/r02c_timeline_ldg.log
timeout 300 python tools/probe_items.py > $O/r02c_items_tma.log 2>&1
cat $O/r02c_timeline_tma.log; cat $O/r02c_timeline_ldg.log; cat $O/r02c_items_tma.log
