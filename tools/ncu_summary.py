#!/usr/bin/env python
"""Summarise an ncu report: per kernel launch the duration, DRAM bytes, achieved DRAM throughput, registers, shared
memory -> JSON on stdout.  usage: ncu_summary.py <report.ncu-rep>"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[col["Kernel Name"]].split("(")[0], "id": r[col["ID"]]}
        for m, name in WANT.items():
            if m in col and r[col[m]] != "":
                v = float(r[col[m]].replace(",", ""))
                u = units[col[m]]
                if u in UNIT_SCALE:
                    v *= UNIT_SCALE[u]
                d[name] = v
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
            if d.get("duration_ns"):
                d["dram_gbs"] = d["dram_bytes"] / d["duration_ns"]
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
