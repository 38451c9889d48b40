#!/bin/bash
tag=${1:-big2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${tag}_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    -k regex:"big_|confusion|gsel|gstats|gflag" python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_list.log 2>&1
grep -c . gpurun_out/${tag}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"big_write|big_load|big_range|big_pass" -s 16 -c 5 -f -o gpurun_out/${tag}_kernels \
    python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
