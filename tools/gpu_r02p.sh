#!/bin/bash
O=gpurun_out; mkdir -p $O
for c in 8 4 2; do
FHSIM_SECTOR_CLUSTER=$c FHSIM_SECTOR_TIMELINE=1 FHSIM_NO_GRAPH=1 timeout 300 python tools/check_sector.py 2 > $O/r02p_timeline_c$c.log 2>&1
echo "== C=$c"; grep "bench cfg3" $O/r02p_timeline_c$c.log | cut -c1-200
grep "sector timeline" $O/r02p_timeline_c$c.log | tail -1 | cut -c1-400
grep "sector timeline" $O/r02p_timeline_c$c.log | tail -1 | grep -o "C[0-9]*\[.*" | cut -c1-700
done
