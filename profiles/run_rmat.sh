#!/bin/bash
# usage: profiles/run_rmat.sh "N list" "exchange list" [workload]   (on the GPU box via gpurun --gpus max(N))
WL=${3:-rmat26}
for n in $1; do for ex in $2; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --workload $WL --steps 5 --warmup 3 2>>gpurun_out/rmat_err.log > gpurun_out/bench_${WL}_${n}gpu_${ex}.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --workload $WL --exchange $ex --steps 5 --warmup 3 2>>gpurun_out/rmat_err.log > gpurun_out/bench_${WL}_${n}gpu_${ex}.json
  fi
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${WL}_${n}gpu_${ex}.json")); c = d["config"]
print("$WL gpus $n exchange $ex: ms/step %.2f  Gedges/s %.2f  local kernel ms (max over ranks) %.2f  xchg in GB (rank0) %.2f"
      % (d["ms_per_step"], d["value"] / 1e9, c["local_kernel_ms_max_over_ranks"], c["exchange_bytes_in_rank0"] / 1e9))
PY
done; done
