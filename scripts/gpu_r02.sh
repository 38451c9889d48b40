#!/bin/bash
# usage: scripts/gpu_r02.sh <tag> [pytest -k expression]
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x ${2:+-k "$2"} > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
cp gpurun_out/image_errors.txt gpurun_out/${tag}_image_errors.txt 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --lookahead 0 --no-extra > gpurun_out/${tag}_bench_seq.json 2> gpurun_out/${tag}_bench_seq.err; echo "bench seq rc=$?"
python bench.py --steps 20 --warmup 5 --lookahead 1 --no-extra > gpurun_out/${tag}_bench_la1.json 2> gpurun_out/${tag}_bench_la1.err; echo "bench la1 rc=$?"
python bench.py --steps 20 --warmup 5 --phase1-stream side --no-extra > gpurun_out/${tag}_bench_side.json 2> gpurun_out/${tag}_bench_side.err; echo "bench side rc=$?"
python bench.py --steps 20 --warmup 5 --phase1-stream side --lookahead 1 --no-extra > gpurun_out/${tag}_bench_side1.json 2> gpurun_out/${tag}_bench_side1.err; echo "bench side1 rc=$?"
python - <<PY
import json
for n in ("bench", "bench_seq", "bench_la1", "bench_side", "bench_side1"):
    try:
        d = json.load(open("gpurun_out/${tag}_%s.json" % n)); r = d["roofline"]
        print(n, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "write", round(r["kernel_ms"], 3), "stats", round(r["stats_kernel_ms"], 3),
              "path", round(r["path"]["frac_create_dataset"], 3), round(r["path"]["frac_with_metrics"], 3), "e2e", round(d["e2e"]["value"], 2),
              "e2e_host", d.get("e2e_host_result"))
        for k, v in d.get("extra", {}).items():
            print("  ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ("workload", "path")}, v.get("path", {}).get("frac_with_metrics"))
    except Exception as e:
        print(n, "ERR", e)
PY
