#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/: one row per launch with
time, tensor-pipe active %, DRAM bytes and achieved GB/s against the measured HBM peak, L2 hit rate, launch geometry.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv out.csv "<command line>" [op names...]
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
COLS = [("time_us", "gpu__time_duration.sum"),
        ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
        ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("cluster_x", "launch__cluster_dim_x"), ("dyn_smem_KB", "launch__shared_mem_per_block_dynamic")]


def main(raw, out, cmd, names):
    peak = 6569.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, metric):
        if metric not in ix or r[ix[metric]] in ("", "n/a"):
            return None
        v = float(r[ix[metric]].replace(",", ""))
        return v * UNIT.get(units[ix[metric]], 1.0)
    with open(out, "w") as f:
        f.write("# %s\n# one row per launch; HBM peak = %.1f GB/s (MEASURED_PEAKS.json); ncu serialises launches and flushes caches,\n"
                "# so times are cold-cache upper bounds -- shares, pipe %% and bytes are what this table is for\n" % (cmd, peak))
        f.write("launch,op,kernel," + ",".join(c[0] for c in COLS) + ",dram_GBps,dram_pct_of_hbm_peak\n")
        for i, r in enumerate(data):
            k = r[ix["Kernel Name"]]
            m = re.match(r"(?:void )?(?:idc::)?([A-Za-z0-9_]+)(<[^>]*>)?", k)
            kn = (m.group(1) + (re.sub(r"\((?:int|bool)\)", "", m.group(2)) if m.group(2) else "")) if m else k[:40]
            vals = [val(r, c[1]) for c in COLS]
            t_us = vals[0] or 0.0
            rd, wr = (vals[2] or 0.0), (vals[3] or 0.0)
            vals[2], vals[3] = rd / 1e6, wr / 1e6
            if vals[12] is not None:
                vals[12] /= 1e3
            gbps = (rd + wr) / 1e9 / (t_us * 1e-6) if t_us else 0.0
            op = names[i] if i < len(names) else ""
            f.write("%d,%s,\"%s\",%s,%.1f,%.2f\n" % (i, op, kn, ",".join("" if v is None else ("%.3f" % v) for v in vals),
                                                   gbps, 100.0 * gbps / peak))
    print("wrote %s (%d launches)" % (out, len(data)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "", sys.argv[4:])
