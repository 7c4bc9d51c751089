#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output into a small JSON (one entry per profiled launch).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/rNN_name.json
"""
import csv
import json
import sys

KEEP = {
    "Kernel Name": "kernel",
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "dram__bytes_read.sum.per_second": "dram_read_TBps",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "lts__t_bytes.sum": "l2_MB",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__cycles_active.avg": "sm_cycles_active_avg",
    "gpc__cycles_elapsed.max": "cycles_elapsed",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
}
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {}
        for k, name in KEEP.items():
            if k in d and d[k] != "":
                try:
                    e[name] = float(d[k].replace(",", ""))
                except ValueError:
                    e[name] = d[k]
                if name in ("dram_read_MB", "dram_write_MB", "l2_MB", "dram_read_TBps", "duration_us"):
                    e[name + "_unit"] = u.get(k, "")
        stalls = {k[len(STALLS):]: float(v) for k, v in d.items()
                  if k.startswith(STALLS) and not k.endswith("_not_issued") and v not in ("", "0")}
        e["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        out.append(e)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
