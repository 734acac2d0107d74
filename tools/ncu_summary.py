#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.json]"""
import csv, json, subprocess, sys, io, re

WANT = {
    "gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "launch__shared_mem_per_block_dynamic": "smem_dyn", "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_bytes.sum": "l2_bytes", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")}
        for k, name in WANT.items():
            if k in idx and r[idx[k]] not in ("", "n/a"):
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                if u in UNIT and ("byte" in u or name == "time_us"):
                    v *= UNIT[u]
                d[name] = v
        if "dram_read" in d:
            d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0)
            d["dram_gbs"] = d["dram_bytes"] / d["time_us"] / 1e3
        res.append(d)
    for d in res:
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()}))
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
