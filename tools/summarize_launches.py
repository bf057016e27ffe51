#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [--md]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mliis::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    return tot, cnt


def main():
    tot, cnt = load(sys.argv[1])
    T = sum(tot.values())
    print("total %.1f us over %d launches (cold-cache, serialised: compare SHARES, not absolutes)\n" % (T, sum(cnt.values())))
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
        print("| `%s` | %d | %.0f | %.1f%% | %.1f |" % (k[:70], cnt[k], v, 100 * v / T, v / cnt[k]))


if __name__ == "__main__":
    main()
