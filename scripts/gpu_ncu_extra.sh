#!/bin/bash
# ncu captures of the remaining kernels on the bench step: confusion_kernel, synth_kernel, rotate_pad / raw gather
tag=${1:-r01x}
out=gpurun_out
mkdir -p $out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"confusion_kernel" -s 3 -c 1 -f -o /tmp/ncu/${tag}_conf \
    python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_conf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"synth_kernel" -c 1 -f -o /tmp/ncu/${tag}_synth \
    python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_synth.log 2>&1
python scripts/ncu_summary.py $out/${tag}_ncu_summary_extra.md /tmp/ncu/${tag}_conf.ncu-rep /tmp/ncu/${tag}_synth.ncu-rep > /dev/null 2> $out/${tag}_summary_extra.err
mv $out/traffic.json $out/${tag}_traffic_extra.json 2>/dev/null
ls -la $out | tail -5
