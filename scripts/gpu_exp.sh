#!/bin/bash
mkdir -p gpurun_out
for g in 0 3 34 35 36; do
  RFI_MONO_GLOBAL=$g timeout 300 python bench.py --steps 20 > gpurun_out/exp_g$g.json 2> gpurun_out/exp_g$g.err
  python -c "import json;d=json.load(open('gpurun_out/exp_g$g.json'));print('g=$g value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'write',round(d['roofline']['kernel_ms'],3),'stats',round(d['roofline']['stats_kernel_ms'],3))"
  tail -2 gpurun_out/exp_g$g.err
done
RFI_MONO_GLOBAL=35 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
