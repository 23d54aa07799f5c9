#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [steps_in_capture] > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print("| kernel | launches/step | ms/step | share |")
    print("|---|---:|---:|---:|")
    for k, v in tot.most_common():
        if v / T < 0.0005:
            continue
        print("| `%s` | %.1f | %.3f | %.1f%% |" % (k, cnt[k] / steps, v / 1e6 / steps, 100 * v / T))
    print("| **total** | %.1f | %.3f | 100%% |" % (sum(cnt.values()) / steps, T / 1e6 / steps))


if __name__ == "__main__":
    main()
