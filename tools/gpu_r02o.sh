#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/check_sector.py 200 > $O/r02o_sector.log 2>&1; echo "rc=$?" >> $O/r02o_sector.log
grep -v "^{" $O/r02o_sector.log | tail -14
FHSIM_SECTOR_TIMELINE=1 FHSIM_NO_GRAPH=1 timeout 300 python tools/check_sector.py 2 > $O/r02o_timeline.log 2>&1
grep "sector timeline" $O/r02o_timeline.log | tail -2 | cut -c1-3000
