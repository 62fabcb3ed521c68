#!/bin/bash
# usage: profiles/ncu_cases.sh tag "case red" ["case red" ...]  → gpurun_out/m_<tag>_<case>_<red>.csv
# targeted metrics (few replays) for the segment-reduce kernels of each case
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,launch__registers_per_thread,launch__grid_size
tag=$1; shift
for cr in "$@"; do
  set -- $cr
  ncu --metrics $M --clock-control none -k regex:"segreduce|segfinish" -c 4 --csv \
      --log-file gpurun_out/m_${tag}_$1_$2.csv python profiles/prof_case.py $1 $2 2 > gpurun_out/m_${tag}_$1_$2.log 2>&1
done
