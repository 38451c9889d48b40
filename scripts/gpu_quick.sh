#!/bin/bash
# quick check: all GPU tests + benches c2 / c3 / c5
tag=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${tag}_tests.log
for w in c2 c3 c5; do
timeout 600 python bench.py --workload $w --steps 10 > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err; echo "$w rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_$w.json'));print('$w value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],3),'write',round(d['roofline']['kernel_ms'],3),'stats',round(d['roofline']['stats_kernel_ms'],3),'frac',round(d['roofline']['frac'],3))"
tail -2 gpurun_out/${tag}_$w.err
done
