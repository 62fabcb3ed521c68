#!/bin/bash
# Runs the reference's own op_bm_scripts UNCHANGED against the shim packages on a B200.
#
# Stage first, in the build container (the reference tree does not exist on the GPU box):
#     bash profiles/run_ref_scripts.sh stage     # copies op_bm_scripts + graph_benchmark into baseline/_ref/
# baseline/_ref/ is git-ignored (never committed) but travels with the gpurun snapshot.  Then:
#     gpurun -- 'bash profiles/run_ref_scripts.sh run'
# Each script writes its own CSV (mem_prof_data/<op>_small.csv, new_data/..., as its source says)
# below gpurun_out/ref_scripts/; stdout/stderr and the wall time of every script are kept beside it.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="$ROOT/baseline/_ref"
case "${1:-run}" in
stage)
  mkdir -p "$REF"
  cp -r /root/reference/op_bm_scripts /root/reference/graph_benchmark "$REF/"
  ls "$REF"
  ;;
run)
  OUT="$ROOT/gpurun_out/ref_scripts"
  mkdir -p "$OUT/mem_prof_data" "$OUT/new_data" "$OUT/datatest" "$OUT/data"
  cd "$OUT" || exit 1
  export PYTHONPATH="$ROOT/gnn-ops-benchmark_b200:$REF"
  shift
  SCRIPTS="${*:-benchmark_scatter_add benchmark_scatter_max benchmark_scatter_min benchmark_scatter_mean benchmark_sparse_coalesce}"
  : > wall_times.txt
  for s in $SCRIPTS; do
    t0=$(date +%s.%N)
    timeout 900 python "$REF/op_bm_scripts/$s.py" > "$s.log" 2>&1
    rc=$?
    t1=$(date +%s.%N)
    echo "$s rc=$rc wall_s=$(python -c "print(round($t1 - $t0, 1))")" | tee -a wall_times.txt
    tail -3 "$s.log"
  done
  python - <<'EOF'
import sys, torch_scatter, torch_sparse
print("torch_scatter ->", torch_scatter.__file__)
print("torch_sparse  ->", torch_sparse.__file__)
EOF
  find . -name "*.csv" | sort
  ;;
esac
