#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into a small tracked text file under profiles/.

    python scripts/ncu_summary.py profiles/r01_summary.md gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]
STALLS = "smsp__average_warps_issue_stalled_"  # prefix of the per-reason stall metrics (ratio per issue)


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    dest, reps = sys.argv[1], sys.argv[2:]
    traffic = {}
    lines = ["# ncu summaries (`ncu --set full --clock-control none`, one launch per kernel, cold-ish cache)", ""]
    for rep in reps:
        hdr, units, launches = raw(rep)
        for r in launches:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            lines.append(f"## {d.get('Kernel Name', '?')}  ({rep.split('/')[-1]})")
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for k in KEYS:
                if k in d and d[k] != "":
                    lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
            st = sorted(((float(v), k[len(STALLS):]) for k, v in d.items()
                         if k.startswith(STALLS) and k.endswith("_per_warp_active.pct") and v not in ("", "n/a")), reverse=True)
            if st:
                lines.append("")
                lines.append("top stall reasons (% of active warps): " +
                             ", ".join(f"{n.replace('_per_warp_active.pct', '')} {v:.1f}" for v, n in st[:6]))
            try:
                rd = float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
                t = float(d["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[u["gpu__time_duration.sum"]]
                lines.append("")
                kname = d.get("Kernel Name", "?").split("<")[0].split("(")[0].replace("void ", "").strip()
                traffic[kname] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                  "duration_ms_under_ncu": t * 1e3, "report": rep.split("/")[-1]}
                lines.append(f"DRAM traffic {rd + wr:.4g} B (read {rd:.4g}, write {wr:.4g}) in {t * 1e3:.3f} ms = "
                             f"{(rd + wr) / t / 1e9:.0f} GB/s under ncu")
            except Exception:
                pass
            lines.append("")
    open(dest, "w").write("\n".join(lines) + "\n")
    if traffic:
        import json
        import os
        tj = os.path.join(os.path.dirname(dest), "traffic.json")
        json.dump(traffic, open(tj, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
