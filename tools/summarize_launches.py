#!/usr/bin/env python
"""Summarise an ncu `--csv` launch list (metrics gpu__time_duration.sum and, when present, dram__bytes_read.sum /
dram__bytes_write.sum) into per-kernel totals (markdown on stdout) and, with --json, the per-launch DRAM traffic of the
tcgen05 GEMM family that bench.py reports as `roofline.traffic`.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [steps_in_capture] [--json profiles/rNN_traffic.json]"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "KB": 1e3, "MB": 1e6, "GB": 1e9, "B": 1.0}
GEMM_FAMILY = ("gemm_tcgen05_kernel", "conv3x3_halo_kernel")


def main():
    args = [a for a in sys.argv[1:]]
    json_out = None
    if "--json" in args:
        i = args.index("--json")
        json_out = args[i + 1]
        del args[i:i + 2]
    path = args[0]
    steps = int(args[1]) if len(args) > 1 else 1
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt, rd, wr = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        m = row.get("Metric Name")
        if m == "gpu__time_duration.sum":
            tot[name] += v
            cnt[name] += 1
        elif m == "dram__bytes_read.sum":
            rd[name] += v
        elif m == "dram__bytes_write.sum":
            wr[name] += v
    T = sum(tot.values())
    have_dram = bool(rd) or bool(wr)
    print("| kernel | launches/step | ms/step | share |" + (" DRAM read GB/step | DRAM write GB/step | GB/s |" if have_dram else ""))
    print("|---|---:|---:|---:|" + ("---:|---:|---:|" if have_dram else ""))
    for k, v in tot.most_common():
        if v / T < 0.0005:
            continue
        line = "| `%s` | %.1f | %.3f | %.1f%% |" % (k, cnt[k] / steps, v / 1e6 / steps, 100 * v / T)
        if have_dram:
            line += " %.3f | %.3f | %.0f |" % (rd[k] / 1e9 / steps, wr[k] / 1e9 / steps, (rd[k] + wr[k]) / v if v else 0.0)
        print(line)
    print("| **total** | %.1f | %.3f | 100%% |" % (sum(cnt.values()) / steps, T / 1e6 / steps)
          + (" %.3f | %.3f | %.0f |" % (sum(rd.values()) / 1e9 / steps, sum(wr.values()) / 1e9 / steps,
                                        (sum(rd.values()) + sum(wr.values())) / T) if have_dram else ""))
    if json_out:
        fam = [k for k in tot if any(g in k for g in GEMM_FAMILY)]
        n = sum(cnt[k] for k in fam)
        out = {"source": path, "steps_in_capture": steps,
               "gemm_family": {"kernels": fam, "launches_per_step": n / steps,
                               "ns_per_step_serialised": sum(tot[k] for k in fam) / steps,
                               "share_of_step_serialised": sum(tot[k] for k in fam) / T,
                               "dram_bytes_per_step": (sum(rd[k] + wr[k] for k in fam) / steps) if have_dram else None,
                               "dram_bytes_per_launch": (sum(rd[k] + wr[k] for k in fam) / n) if (have_dram and n) else None},
               "all_kernels": {"launches_per_step": sum(cnt.values()) / steps, "ns_per_step_serialised": T / steps,
                               "dram_bytes_per_step": ((sum(rd.values()) + sum(wr.values())) / steps) if have_dram else None}}
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
