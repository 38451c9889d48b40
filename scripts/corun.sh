#!/bin/bash
# co-residency experiment: writer limited to 1 CTA / SM by a padded shared-memory request, statistics kernel at 45 KB
run() { tag=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --no-extra ${EXTRA} > gpurun_out/corun_$tag.json 2> gpurun_out/corun_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/corun_$tag.json")); r = d["roofline"]
print("$tag", "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "write", round(r["kernel_ms"], 3), "stats", round(r["stats_kernel_ms"], 3))
PY
}
EXTRA="" run base A=1
EXTRA="" run gs3 RFI_MONO_GS=3
EXTRA="" run wpad5 RFI_WRITER_SMEM_PAD_KB=5
EXTRA="" run stats2 RFI_STATS_SMEM_PAD_KB=20
EXTRA="--phase1-stream side" run side_gs3_wpad5 RFI_MONO_GS=3 RFI_WRITER_SMEM_PAD_KB=5
EXTRA="--phase1-stream side" run side_gs3 RFI_MONO_GS=3
EXTRA="--phase1-stream side --lookahead 1" run side1_gs3_wpad5 RFI_MONO_GS=3 RFI_WRITER_SMEM_PAD_KB=5
