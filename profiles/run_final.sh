#!/bin/bash
# Round-end check on one B200: GPU test suite, smoke(), the driver's bench line, the ncu launch list of the
# same command restricted to this library's kernels.  usage (gpurun): bash profiles/run_final.sh TAG
T=${1:-final}; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/${T}_bench_1gpu.json"))
print("ms/step", d["ms_per_step"], "G edges/s", d["value"] / 1e9, "parity", d["parity"]["max_err_over_bound"], d["parity"]["ok"],
      "roofline", {k: d["roofline"][k] for k in ("frac", "dram_frac", "kernel_ms")}, "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in d.get("per_config", {}).items():
    print("   %8.3f ms  %.3f %s  %s" % (v["ms"], v["frac_of_peak"], ("dram %.2f" % v["frac_dram"]) if "frac_dram" in v else "         ", k))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'seg|radix|scan|plan|permute|narrow|lists|expand|push|scatter_' -c 300 --csv \
  --log-file $O/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --per-config 0 > $O/${T}_launches_bench.log 2>&1
echo "ncu rc=$?"
