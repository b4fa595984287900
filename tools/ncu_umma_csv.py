#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` of the 25 umma_conv launches of one bench step into the
per-layer table kept under profiles/ (usage: ncu_umma_csv.py raw.csv out.csv "<command line>")."""
import csv
import re
import sys

OPS = ["c1_2", "c2_1", "c2_2", "c3_1", "c3_2", "c3_3", "c4_1", "c4_2", "c4_3", "c5_1", "c5_2", "c5_3", "c6_1", "c6_2",
       "c6_3", "c7_1", "c7_2", "c7_3", "up8", "c8_2", "c8_3", "up9", "c9_2", "up10", "c10_2"]
COLS = [("time[ms]", "gpu__time_duration.sum", 1e-6),                     # ns -> ms (unit checked below)
        ("tensor_pipe_active_pct[%]", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("dram_read[Gbyte]", "dram__bytes_read.sum", None),
        ("dram_write[Gbyte]", "dram__bytes_write.sum", None),
        ("l2_hit_pct[%]", "lts__t_sector_hit_rate.pct", 1),
        ("l2_throughput_pct[%]", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("regs[register/thread]", "launch__registers_per_thread", 1),
        ("cluster_x[]", "launch__cluster_dim_x", 1),
        ("cycles[cycle]", "sm__cycles_elapsed.max", 1)]
UNIT = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(raw, out, cmd):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    kname = ix["Kernel Name"]
    with open(out, "w") as f:
        f.write("# %s\n# one row per conv launch of one step; units as reported by ncu\n" % cmd)
        f.write("op,template<BN MT CG SPLIT>," + ",".join(c[0] for c in COLS) + "\n")
        for op, r in zip(OPS, data):
            m = re.search(r"umma_conv_kernel<([^>]*)>", r[kname])
            tpl = re.sub(r"\((?:int|bool)\)", "", m.group(1)) if m else "?"
            vals = []
            for label, metric, scale in COLS:
                if metric not in ix:
                    vals.append("nan")
                    continue
                v = float(r[ix[metric]].replace(",", "") or "nan")
                u = units[ix[metric]]
                if scale is None or label.startswith("time"):
                    v *= UNIT.get(u, 1.0)
                vals.append("%.6f" % v if label.endswith("]") and "regs" not in label and "cluster" not in label else "%d" % v)
            f.write('%s,"<%s>",%s\n' % (op, tpl, ",".join(vals)))
    if len(data) != len(OPS):
        print("warning: %d launches in the capture, expected %d" % (len(data), len(OPS)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
