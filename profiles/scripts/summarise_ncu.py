#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into one JSON record per kernel launch with the metrics the roofline
discussion needs.  usage: summarise_ncu.py report.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6,
              "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")][:80]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                rec[name] = v * UNIT_SCALE.get(units[i].split("/")[0], 1.0) if ("bytes" in name or name == "duration_us") else v
        if "dram_read_bytes" in rec and "duration_us" in rec:
            rec["dram_traffic_bytes"] = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
            rec["dram_gbs"] = rec["dram_traffic_bytes"] / rec["duration_us"] / 1e3
        out.append(rec)
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js + "\n")
    print(js)


if __name__ == "__main__":
    main()
