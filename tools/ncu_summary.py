#!/usr/bin/env python
"""ncu report -> one JSON summary per captured launch (the counters the roofline arguments use), for profiles/.
usage: ncu_summary.py report.ncu-rep [points-per-launch ...]   (points: optional, in launch order, for per-point figures)"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
        "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    rep = sys.argv[1]
    pts = [float(x) for x in sys.argv[2:]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for li, r in enumerate(rows[2:]):
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        stalls = {}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                if u in UNIT and ("bytes" in h or "time" in h):
                    x *= UNIT[u]
                d[h] = x
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
                except ValueError:
                    pass
        if stalls:
            tot = sum(stalls.values()) or 1.0
            d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        if li < len(pts):
            d["points"] = pts[li]
            if "smsp__inst_executed.sum" in d:
                d["warp_inst_per_point"] = d["smsp__inst_executed.sum"] / pts[li]
                d["thread_inst_per_point"] = 32 * d["warp_inst_per_point"]
            d["dram_bytes_per_point"] = (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / pts[li]
        res.append(d)
    print(json.dumps(res if len(res) != 1 else res[0], indent=1))


if __name__ == "__main__":
    main()
