#!/bin/bash
# usage: profiles/ncu_dur.sh tag case red   → gpurun_out/dur_<tag>.csv with per-launch duration / dram / inst
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
python profiles/prof_case.py $2 $3 2 > gpurun_out/pc_$1.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:"segreduce|segfinish" --csv --log-file gpurun_out/dur_$1.csv python profiles/prof_case.py $2 $3 2 > gpurun_out/pc2_$1.log 2>&1
