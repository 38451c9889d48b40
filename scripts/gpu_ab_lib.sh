#!/bin/bash
# A/B of two builds of the library on the full bench line (RFI_B200_LIB selects the build): old, new, old, new
old=${1:-gpurun_ab/lib_old.so}
mkdir -p gpurun_out
for i in 1 2; do
  RFI_B200_LIB=$PWD/$old timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/ab_old$i.json 2> gpurun_out/ab_old$i.err
  timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/ab_new$i.json 2> gpurun_out/ab_new$i.err
done
python - <<PY
import json
for t in ("old1", "new1", "old2", "new2"):
    d = json.load(open(f"gpurun_out/ab_{t}.json")); r = d["roofline"]; e = d["extra"]
    print(t, "c2 %.3f ms (stats %.3f write %.3f)" % (d["ms_per_step"], r["stats_kernel_ms"], r["kernel_ms"]),
          "| c3 %.2f (p1 %.2f w %.2f)" % (e["c3_shard"]["ms_per_step"], e["c3_shard"]["phase1_ms"], e["c3_shard"]["writer_ms"]),
          "| c5 %.2f (p1 %.2f w %.2f)" % (e["c5_chunk"]["ms_per_step"], e["c5_chunk"]["phase1_ms"], e["c5_chunk"]["writer_ms"]))
PY
