#!/bin/bash
# ncu captures of the dominant kernels, summarised ON the box (reports are too large to bring back):
# launch lists, ncu_summary.md + traffic.json, per-source-line instruction counts.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out /tmp/ncu
lib=rfi_toolbox_b200/_lib/librfi_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches.csv \
    -k regex:"tile_stats|write_patches|confusion|flags_count" python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $out/${tag}_launches_c5.csv \
    -k regex:"big_|confusion|gsel|gstats|gflag" python bench.py --workload c5 --steps 2 --warmup 3 > $out/${tag}_ncu_list_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:write_patches -s 3 -c 1 -f -o /tmp/ncu/${tag}_write \
    python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_write.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_stats_mono -s 3 -c 1 -f -o /tmp/ncu/${tag}_stats \
    python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_stats.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"big_" -s 40 -c 8 -f -o /tmp/ncu/${tag}_big \
    python bench.py --workload c5 --steps 1 --warmup 3 > $out/${tag}_ncu_big.log 2>&1
python scripts/ncu_summary.py $out/${tag}_ncu_summary.md /tmp/ncu/${tag}_write.ncu-rep /tmp/ncu/${tag}_stats.ncu-rep /tmp/ncu/${tag}_big.ncu-rep > /dev/null 2> $out/${tag}_summary.err
python scripts/ncu_lines.py /tmp/ncu/${tag}_stats.ncu-rep $lib tile_stats_mono 60 > $out/${tag}_stats_mono_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu/${tag}_write.ncu-rep $lib write_patches 40 > $out/${tag}_write_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu/${tag}_big.ncu-rep $lib big_write 40 > $out/${tag}_big_write_lines.txt 2>&1
ls -la /tmp/ncu $out
