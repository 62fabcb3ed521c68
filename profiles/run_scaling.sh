#!/bin/bash
# usage: profiles/run_scaling.sh N "stages list"   (run on the GPU box via gpurun --gpus N)
N=$1; shift
for st in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + st)) \
    bench.py --gpus $N --steps 10 --warmup 3 --stages $st 2>>gpurun_out/scaling_err.log > gpurun_out/bench_${N}gpu_st${st}.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu_st${st}.json"))
print("gpus $N stages $st ms/step %.3f  Gedges/s %.2f  e2e ms %.1f" % (d["ms_per_step"], d["value"] / 1e9, d["e2e"]["ms_per_step"]))
PY
done
