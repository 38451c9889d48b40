#!/bin/bash
tag=${1:-big3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bigtile.py tests/test_gpu_generic.py -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests(cluster) rc=$?"
tail -4 gpurun_out/${tag}_tests.log
RFI_BIG_NO_CLUSTER=1 timeout 900 python -m pytest tests/test_gpu_bigtile.py -q -x > gpurun_out/${tag}_tests_nc.log 2>&1; echo "tests(no cluster) rc=$?"
tail -3 gpurun_out/${tag}_tests_nc.log
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_c5.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['stats_kernel_ms'],d['roofline']['frac'])"
tail -3 gpurun_out/${tag}_c5.err
RFI_BIG_NO_CLUSTER=1 timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/${tag}_c5_nc.json 2> gpurun_out/${tag}_c5_nc.err; echo "bench(nc) rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_c5_nc.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['stats_kernel_ms'],d['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/${tag}_launches.csv \
    -k regex:"big_|confusion|gsel|gstats|gflag" python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_list.log 2>&1
grep -c . gpurun_out/${tag}_launches.csv
