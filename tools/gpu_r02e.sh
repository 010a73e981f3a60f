#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/probe_timeline.py --rebuild > $O/r02e_timeline_tma.log 2>&1; echo "rc=$?" >> $O/r02e_timeline_tma.log
grep -v -i "warning\|remark\|^ *\^\|^$\|static int\|double theta" $O/r02e_timeline_tma.log | cut -c1-420 | tail -16
