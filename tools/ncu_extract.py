#!/usr/bin/env python3
"""Selected counters of every kernel of an `ncu --page raw --csv` export, as JSON (dev tool; the summaries under profiles/ are
written from its output):  tools/ncu_extract.py raw.csv [raw2.csv ...]"""
import csv
import json
import sys

KEYS = {"gpu__time_duration.sum": "duration", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
        "launch__registers_per_thread": "registers", "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instr",
        "smsp__inst_executed.sum": "warp_instructions", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_lsu_data_pipe_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_wavefronts_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__shared_mem_per_block_static": "static_smem", "launch__shared_mem_per_block_dynamic": "dynamic_smem"}
out = []
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        d = {"kernel": r[ix["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", ""), "report": path.split("/")[-1]}
        for k, name in KEYS.items():
            if k in ix:
                try:
                    d[name] = float(r[ix[k]])
                    if units[ix[k]] not in ("", "%"):
                        d[name + "_unit"] = units[ix[k]]
                except ValueError:
                    pass
        st = []
        for h in hdr:
            if "issue_stalled" in h and "ratio" in h and "not_issued" not in h:
                try:
                    st.append((float(r[ix[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        d["top_stalls"] = {h: round(v, 2) for v, h in sorted(st, reverse=True)[:5]}
        out.append(d)
print(json.dumps(out, indent=1))
