#!/bin/bash
# usage: scripts/gpu_fused.sh <tag>   -- tests, then bench A/B of the metrics stream / single launch
tag=${1:-fused}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_single_launch.py tests/test_gpu_pairs.py -q -x > gpurun_out/${tag}_new_tests.log 2>&1; echo "new tests rc=$?"; tail -25 gpurun_out/${tag}_new_tests.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${tag}_tests.log
cp gpurun_out/image_errors.txt gpurun_out/${tag}_image_errors.txt 2>/dev/null
run() { n=$1; shift
  timeout 600 env $ENVV python bench.py --steps 30 --warmup 5 --no-extra "$@" > gpurun_out/${tag}_$n.json 2> gpurun_out/${tag}_$n.err; echo "$n rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$n.json")); r = d["roofline"]
    print("$n", "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), r["kernel"], round(r["kernel_ms"], 3), "frac", round(r["frac"], 3),
          "stats", r.get("stats_kernel_ms"), "path", round(r["path"]["frac_create_dataset"], 3), round(r["path"]["frac_with_metrics"], 3), "e2e", round(d["e2e"]["value"], 2))
except Exception as e:
    print("$n", "ERR", e)
PY
}
ENVV="A=1" run side8
ENVV="RFI_CONFUSION_CTAS_PER_SM=2" run side2
ENVV="RFI_CONFUSION_CTAS_PER_SM=1" run side1
ENVV="RFI_CONFUSION_CTAS_PER_SM=4" run side4
ENVV="A=1" run main8 --metrics-stream main
ENVV="RFI_CONFUSION_CTAS_PER_SM=2" run main2 --metrics-stream main
ENVV="A=1" run single --single-launch
