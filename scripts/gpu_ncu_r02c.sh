#!/bin/bash
# end-of-round-2 captures of the two writers after their plain variant (see scripts/gpu_ncu_r02.sh for the full set)
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out /tmp/ncu
lib=rfi_toolbox_b200/_lib/librfi_b200.so
cap() { name=$1; regex=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -f -o /tmp/ncu/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1; }
cap write write_patches 3 python bench.py --steps 1 --warmup 3 --no-extra
cap bigwrite big_write 2 python bench.py --workload c5 --steps 1 --warmup 2 --no-extra
reps=""
for n in write bigwrite; do [ -f /tmp/ncu/${tag}_$n.ncu-rep ] && reps="$reps /tmp/ncu/${tag}_$n.ncu-rep"; done
python scripts/ncu_summary.py $out/${tag}_ncu_summary.md $reps > /dev/null 2> $out/${tag}_summary.err
mv $out/traffic.json $out/${tag}_traffic.json 2>/dev/null
python scripts/ncu_lines.py /tmp/ncu/${tag}_write.ncu-rep $lib write_patches 40 > $out/${tag}_write_lines.txt 2>&1
ls -la $out | tail -6
