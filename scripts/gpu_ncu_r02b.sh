#!/bin/bash
# end-of-round-2 ncu evidence for the kernels that changed after scripts/gpu_ncu_r02.sh ran (phase 1 and the
# pair sweep): launch list of the bench step, `--set full` captures, per-source-line instruction counts.
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out /tmp/ncu
lib=rfi_toolbox_b200/_lib/librfi_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_launches.csv \
    -k regex:"tile_stats|write_patches|confusion|flags_count" python bench.py --steps 3 --warmup 3 --no-extra > $out/${tag}_ncu_list.log 2>&1
cap() { name=$1; regex=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -f -o /tmp/ncu/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1; }
cap stats tile_stats_mono 3 python bench.py --steps 1 --warmup 3 --no-extra
cap statsc3 tile_stats_mono 2 python bench.py --workload c3 --baselines 8 --steps 1 --warmup 2 --no-extra
cap pairs pair_sweep 1 python scripts/time_pairs.py 20000 1
reps=""
for n in stats statsc3 pairs; do [ -f /tmp/ncu/${tag}_$n.ncu-rep ] && reps="$reps /tmp/ncu/${tag}_$n.ncu-rep"; done
python scripts/ncu_summary.py $out/${tag}_ncu_summary.md $reps > /dev/null 2> $out/${tag}_summary.err
mv $out/traffic.json $out/${tag}_traffic.json 2>/dev/null
python scripts/ncu_lines.py /tmp/ncu/${tag}_stats.ncu-rep $lib tile_stats_mono 60 > $out/${tag}_stats_mono_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu/${tag}_pairs.ncu-rep $lib pair_sweep 50 > $out/${tag}_pair_sweep_lines.txt 2>&1
ls -la $out | tail -12
