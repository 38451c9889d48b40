#!/bin/bash
# usage: scripts/gpu_final.sh <tag>   -- smoke(), the whole GPU suite, the full bench line and the reference arm
tag=${1:-final}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 300 python scripts/time_pairs.py 100000 5 > gpurun_out/${tag}_pairs.log 2>&1; tail -2 gpurun_out/${tag}_pairs.log
bash scripts/gpu_full.sh ${tag}
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/${tag}_ref.json
