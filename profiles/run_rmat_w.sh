#!/bin/bash
# usage: profiles/run_rmat_w.sh N "row-weight list"
n=$1
for w in $2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + w)) \
    bench.py --gpus $n --workload rmat26 --exchange needed --row-weight $w --steps 5 --warmup 3 2>>gpurun_out/rmat_err.log > gpurun_out/bench_rmat26_${n}gpu_w${w}.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_rmat26_${n}gpu_w${w}.json")); c = d["config"]
print("w $w: ms/step %.2f Gedges/s %.2f kernel max %.2f | rank0 kernel %.2f edges %d rows %d"
      % (d["ms_per_step"], d["value"] / 1e9, c["local_kernel_ms_max_over_ranks"], d["roofline"]["kernel_ms"], c["edges_rank0"], c["rows_rank0"]))
PY
done
