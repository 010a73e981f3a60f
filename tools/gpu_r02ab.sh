#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02ab_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-regime > $O/r02ab_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02ab_launches.csv")) if len(r)>5 and r[0].isdigit()]
# find the last occurrence of k_sector_pool and print the 20 launches before it
names=[(r[4].split("(")[0][-40:], float(r[-1].replace(",",""))) for r in rows]
idx=[i for i,(n,_) in enumerate(names) if "k_sector_pool" in n]
if idx:
    i=idx[len(idx)//2]
    tot=0
    for n,t in names[i-18:i+1]:
        print(f"{n:42s} {t/1000:8.2f} us"); tot+=t
    print("sum", tot/1000)
PY
