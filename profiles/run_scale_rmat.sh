#!/bin/bash
# usage (gpurun --gpus N): bash profiles/run_scale_rmat.sh N TAG [extra bench.py flags]
# bench.py exactly as the driver launches it at N GPUs; one JSON line into gpurun_out/<TAG>_bench_<N>gpu.json
N=$1; TAG=$2; shift 2
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 3 "$@" > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
fi
echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
print("N=%d ms/step %.3f  G edges/s %.2f  parity %s  clocks %s  e2e %.3g  launches %s" % (
    d["n_gpus"], d["ms_per_step"], d["value"] / 1e9, d["parity"].get("max_err_over_bound"),
    d["clocks"].get("sm_mhz"), d["e2e"]["value"], d["gpu_launches"]))
print("   detail", {k: d["detail"][k] for k in ("stages", "stage_fracs", "split", "push_chunk", "local_reduce_ms_max_over_ranks", "exchange_only_ms_max_over_ranks")})
PY
