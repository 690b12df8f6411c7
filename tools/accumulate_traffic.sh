#!/bin/bash
# dram bytes per launch of the accumulate kernel inside one MSM of 2^k points, k = 24, 23, 22, 21 (profiles/r02_accumulate_traffic.json)
for k in 24 23 22 21; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:msm_accumulate --csv --log-file /tmp/at.csv python tools/prof_r2.py msm24 $k > /dev/null 2>&1
  python - $k <<PY
import csv,sys
rows=list(csv.reader(open("/tmp/at.csv",errors="replace")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hh=rows[h]
v={r[hh.index("Metric Name")]:float(r[hh.index("Metric Value")].replace(",","")) for r in rows[h+1:] if len(r)>5}
print(sys.argv[1], v["dram__bytes_read.sum"]+v["dram__bytes_write.sum"], v["dram__bytes_read.sum"], v["dram__bytes_write.sum"], v["gpu__time_duration.sum"])
PY
done
