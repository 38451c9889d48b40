#!/bin/bash
# quick check: metric tests + parity tests + default bench (+ c5 short)
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'write',d['roofline']['kernel_ms'],'stats',d['roofline']['stats_kernel_ms'],'frac',d['roofline']['frac'])"
tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --workload c5 --steps 5 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err; echo "c5 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_c5.json'));print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'write',d['roofline']['kernel_ms'],'stats',d['roofline']['stats_kernel_ms'],'frac',d['roofline']['frac'])"
