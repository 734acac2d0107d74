#!/usr/bin/env python3
"""Summarise ncu reports on the GPU box: ncu_extract.py <out.json> <name=report.ncu-rep> ...  -> per kernel: duration, DRAM bytes, lanes per
instruction, issue utilisation, occupancy, registers, top stall reasons (all from `ncu --page raw --csv`)."""
import csv, io, json, subprocess, sys
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        out[name] = {"error": "empty report"}
        continue
    hdr, units = rows[0], rows[1]
    ks = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        def f(k, scale=1.0):
            try:
                return float(d[k].replace(",", "")) * scale
            except Exception:
                return None
        def bytes_of(k):
            v = f(k)
            if v is None:
                return None
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u.get(k, "byte"), 1)
        dur = f("gpu__time_duration.sum")
        dur_ms = None if dur is None else dur * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u.get("gpu__time_duration.sum", "ns"), 1e-6)
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        ti, wi = f("smsp__thread_inst_executed.sum"), f("smsp__inst_executed.sum")
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: f(k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
        top = dict(sorted(((k, v) for k, v in stalls.items() if v), key=lambda kv: -kv[1])[:5])
        ks.append({"kernel": d.get("Kernel Name"), "duration_ms": dur_ms, "dram_bytes": None if rd is None else rd + wr, "dram_read": rd, "dram_write": wr,
                   "dram_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "lanes_per_instruction": None if not wi else ti / wi,
                   "issue_slots_busy_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active") or f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "achieved_occupancy_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"), "registers": f("launch__registers_per_thread"),
                   "warp_instructions": wi, "l2_hit_pct": f("lts__t_sector_hit_rate.pct"), "stall_cycles_per_issue": top})
    out[name] = ks if len(ks) > 1 else ks[0]
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: ({kk: v.get(kk) for kk in ("duration_ms", "dram_bytes", "lanes_per_instruction", "registers")} if isinstance(v, dict) else len(v)) for k, v in out.items()}))
