#!/bin/bash
# per-SASS-address executed counts of the phase-1 kernel (to attribute subroutines that carry no line info)
out=gpurun_out; mkdir -p $out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_stats_mono -s 3 -c 1 -f -o /tmp/ncu/addr python bench.py --steps 1 --warmup 3 --no-extra > $out/addr_ncu.log 2>&1
ncu -i /tmp/ncu/addr.ncu-rep --page source --csv 2>/dev/null | python -c "
import csv, sys
r = csv.reader(sys.stdin); w = csv.writer(sys.stdout); idx = None
for row in r:
    if 'Address' in row and 'Instructions Executed' in row:
        idx = [row.index(c) for c in ('Address', 'Source', 'Instructions Executed', '# Samples')]
    if idx and len(row) > max(idx): w.writerow([row[i] for i in idx])
" | gzip > $out/addr_stats.csv.gz
ls -la $out/addr_stats.csv.gz
