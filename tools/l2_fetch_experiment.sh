#!/bin/bash
# DRAM traffic and duration of the accumulate kernel of one 2^24 MSM under cudaLimitMaxL2FetchGranularity = default / 32 / 64 / 128
for g in default 32 64 128; do
  if [ $g = default ]; then unset ZKB_L2_FETCH_GRANULARITY; else export ZKB_L2_FETCH_GRANULARITY=$g; fi
  echo "== granularity $g"
  python tools/prof_r2.py msm24 24 2>&1 | tail -1
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:msm_accumulate --csv --log-file /tmp/fg.csv python tools/prof_r2.py msm24 24 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("/tmp/fg.csv",errors="replace")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hh=rows[h]
for r in rows[h+1:]:
    if len(r)>5: print("  ", r[hh.index("Metric Name")], r[hh.index("Metric Unit")], r[hh.index("Metric Value")])
PY
done
