#!/bin/bash
tag=${1:-e2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'write',d['roofline']['kernel_ms'],'stats',d['roofline']['stats_kernel_ms'],'frac',d['roofline']['frac'])"
tail -3 gpurun_out/${tag}_bench.err
