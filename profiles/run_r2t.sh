#!/bin/bash
# Evidence run (1 GPU): GPU test suite, the per-config op table with same-box native torch timings, the
# driver's bench line, ncu DRAM bytes + L2 hit rate per config, full captures of the planned scatter and the
# radix scatter.  usage (gpurun): bash profiles/run_r2t.sh
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2t_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/r2t_pytest.txt
python profiles/bench_ops.py > $O/r2s_ops_table.jsonl 2> $O/r2s_ops.err; echo "ops rc=$?"
python bench.py > $O/r2t_bench_1gpu.json 2> $O/r2t_bench_1gpu.err; echo "bench rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for c in "reddit max" "reddit sum" "reddit_bf16 max" "reddit_bf16 sum" "c4spmm sum" "c5 sum" "rmat26 sum" "products sum" "c1 sum"; do
  set -- $c
  timeout 900 ncu --metrics $M --clock-control none -k regex:'segreduce|segfinish' --csv \
    --log-file $O/r2t_ncu_$1_$2.csv python profiles/prof_case.py $1 $2 1 > $O/r2t_ncu_$1_$2.log 2>&1
  echo "ncu $1 $2 rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_planned -c 8 -f -o $O/r2t_planned \
  python profiles/prof_fullshape.py > $O/r2t_planned.log 2>&1; echo "ncu planned rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -c 2 -f -o $O/r2t_sort \
  python profiles/prof_sort.py 268435456 32 > $O/r2t_sort.log 2>&1; echo "ncu sort rc=$?"
