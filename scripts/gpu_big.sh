#!/bin/bash
# GPU visit for the big-tile path: its parity tests, the generic-path tests, a c5 bench, launch list.
tag=${1:-big}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bigtile.py tests/test_gpu_generic.py -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err; echo "bench rc=$?"
cat gpurun_out/${tag}_c5.json; tail -5 gpurun_out/${tag}_c5.err
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --workload c5 --baselines 2 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_list.log 2>&1
grep -c . gpurun_out/${tag}_launches.csv
fi
