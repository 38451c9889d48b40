#!/bin/bash
# One GPU visit: tests, benches (c2 = the bench workload, c3, c5), host profile, ncu launch lists,
# ncu full captures of the dominant kernels of both paths.
# usage: scripts/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
python bench.py --workload c5 --steps 10 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; echo "bench c5 rc=$?"
cat gpurun_out/${tag}_bench_c5.json
python bench.py --workload c3 --steps 5 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench c3 rc=$?"
cat gpurun_out/${tag}_bench_c3.json
python scripts/hostprof.py > gpurun_out/${tag}_hostprof.log 2>&1
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    -k regex:"tile_stats|write_patches|confusion|flags_count" python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches_c5.csv \
    -k regex:"big_|confusion|gsel|gstats|gflag" python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_list_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:write_patches -s 3 -c 1 -f -o gpurun_out/${tag}_write \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_write.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_stats_mono -s 3 -c 1 -f -o gpurun_out/${tag}_stats \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_stats.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"big_write|big_load|big_pass|big_count" -s 15 -c 5 -f -o gpurun_out/${tag}_big \
    python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_big.log 2>&1
fi
cp rfi_toolbox_b200/_lib/librfi_b200.so gpurun_out/${tag}_lib.so
ls -la gpurun_out | tail -12
