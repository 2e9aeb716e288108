#!/usr/bin/env python3
"""Summarises an ncu `--csv --log-file` launch list (profiles/run_ncu_r02.sh) per kernel: launches, device time, DRAM bytes and -- when
the counters are there -- L2 hit rate and the sectors fetched per global load request (32 for a fully coalesced 128-byte request of a
warp, up to 32 x 1 for 32 scattered 32-byte sectors; the engine's table probes are one sector per lane by design).

    python profiles/summarize_ncu.py gpurun_out/r02_steady.csv [--json profiles/r02_traffic.json] > profiles/r02_steady_block_summary.md
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    launches = OrderedDict()
    for r in rows:
        k = r["ID"]
        e = launches.setdefault(k, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        v = r["Metric Value"].replace(",", "")
        try:
            v = float(v)
        except ValueError:
            continue
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        e[r["Metric Name"]] = v
    return list(launches.values())


def short(name):
    m = re.search(r"(?:fqsk::)?(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name.split("(")[0][:60]


def main():
    path = sys.argv[1]
    L = load(path)
    agg = OrderedDict()
    for e in L:
        a = agg.setdefault(short(e["name"]), defaultdict(float))
        a["n"] += 1
        for k, v in e.items():
            if isinstance(v, float):
                a[k] += v
    tot_us = sum(a["gpu__time_duration.sum"] for a in agg.values())
    tot_rd = sum(a["dram__bytes_read.sum"] for a in agg.values())
    tot_wr = sum(a["dram__bytes_write.sum"] for a in agg.values())
    print(f"{len(L)} launches, {tot_us / 1e3:.3f} ms of kernel time, DRAM {tot_rd / 1e9:.3f} GB read + {tot_wr / 1e9:.3f} GB written = {(tot_rd + tot_wr) / 1e9:.3f} GB\n")
    have_l1 = any("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum" in a for a in agg.values())
    hdr = "| us | share | launches | DRAM read MB | DRAM write MB | DRAM GB/s |" + (" L2 hit % | sectors / ld request | sectors / st request | warps active % |" if have_l1 else "") + " kernel |"
    print(hdr)
    print("|" + "---:|" * (hdr.count("|") - 2) + "---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        us = a["gpu__time_duration.sum"]
        rd, wr = a["dram__bytes_read.sum"], a["dram__bytes_write.sum"]
        row = f"| {us:.1f} | {100 * us / max(tot_us, 1e-9):.1f}% | {int(a['n'])} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / max(us, 1e-9) / 1e3:.0f} |"
        if have_l1:
            n = a["n"]
            ldq, lds = a["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"], a["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
            stq, sts = a["l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"], a["l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
            row += f" {a['lts__t_sector_hit_rate.pct'] / n:.1f} | {lds / ldq if ldq else 0:.2f} | {sts / stq if stq else 0:.2f} | {a['sm__warps_active.avg.pct_of_peak_sustained_active'] / n:.1f} |"
        print(row + f" `{name}` |")
    if "--json" in sys.argv:
        out = sys.argv[sys.argv.index("--json") + 1]
        json.dump({"dram_bytes_per_steady_step": tot_rd + tot_wr, "dram_read_bytes": tot_rd, "dram_write_bytes": tot_wr, "kernel_us": tot_us, "launches": len(L),
                   "note": f"ncu launch list of ONE steady-state block (one 51 k-read segment + its sync): {path}; summed over its {len(L)} launches (profiles/run_ncu_r02.sh)"},
                  open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
