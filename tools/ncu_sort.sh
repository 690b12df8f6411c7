set -e
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'sort_pass_kernel|msm_entry_pass0_kernel|msm_entry_hist_kernel' -c 4 -o /tmp/sortrep python tools/prof_r2.py msm24 24 > gpurun_out/ncu_sort.log 2>&1
ncu -i /tmp/sortrep.ncu-rep --page raw --csv > gpurun_out/ncu_sort_raw.csv
ncu -i /tmp/sortrep.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_sort_src.csv 2>/dev/null || true
ls -la gpurun_out/ncu_sort*
