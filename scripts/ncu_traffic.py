#!/usr/bin/env python
"""Extract per-launch DRAM traffic and a few headline metrics of the captured kernels from `ncu --set full` reports
(gpurun_out/prof_*.ncu-rep) into profiles/<round>_ncu_summary.json, and the dominant kernel's figure into
profiles/conv_traffic.json (read by bench.py for roofline.traffic).  Usage: ncu_traffic.py <round tag> rep [rep ...]"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    # the gather's real ceiling: wavefronts through the L1 data pipe (one per distinct line of a warp request) and the
    # register write-back of the loaded rows
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_lsu_wavefronts_pct",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "l1_lsu_writeback_active_pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_load_requests",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_load_sectors",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, n in enumerate(hdr):
            if n in WANT and r[i] != "":
                v = float(r[i].replace(",", ""))
                d[WANT[n]] = v * UNIT.get(units[i], 1.0)
        res.append(d)
    return res


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    allk = []
    for rep in reps:
        allk += parse(rep)
    json.dump(allk, open(os.path.join(root, "profiles", "%s_ncu_summary.json" % tag), "w"), indent=1)
    conv = [k for k in allk if "spconv_ts_kernel" in k["kernel"] or "spconv_tr_kernel" in k["kernel"]]
    if conv:
        t = sum(k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0) for k in conv) / len(conv)
        json.dump({"dram_bytes_per_launch": t, "launches_captured": len(conv),
                   "source": "ncu --set full, one launch each of " + ", ".join(sorted({k["kernel"].split("(")[0] for k in conv})),
                   "note": "cold-cache capture (ncu flushes caches between replays): an upper bound of the in-step traffic"},
                  open(os.path.join(root, "profiles", "conv_traffic.json"), "w"), indent=1)
    for k in allk:
        print(k["kernel"][:70], {a: round(b, 1) for a, b in k.items() if a != "kernel"})


if __name__ == "__main__":
    main()
