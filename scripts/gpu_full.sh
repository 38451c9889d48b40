#!/bin/bash
# usage: scripts/gpu_full.sh <tag>   -- the whole GPU suite, then the full bench line (extras included)
tag=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/${tag}_tests.log
cp gpurun_out/image_errors.txt gpurun_out/${tag}_image_errors.txt 2>/dev/null
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json")); r = d["roofline"]
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), r["kernel"], round(r["kernel_ms"], 3), "frac", round(r["frac"], 3), "stats", r.get("stats_kernel_ms"),
      "path", round(r["path"]["frac_create_dataset"], 3), round(r["path"]["frac_with_metrics"], 3), "e2e", round(d["e2e"]["value"], 2), "e2e_host", (d.get("e2e_host_result") or {}).get("value"))
for k, v in d.get("extra", {}).items():
    print("  ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ("workload", "path")}, v.get("path", {}).get("frac_with_metrics"))
PY
