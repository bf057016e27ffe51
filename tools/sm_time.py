#!/usr/bin/env python
"""Which kernels consume the machine?  Sums `sm__cycles_active.sum` (SM-cycles a kernel keeps SMs busy) and the
duration per kernel name from an ncu csv:

    ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,launch__grid_size --clock-control none --csv \
        --log-file gpurun_out/smtime.csv python tools/prof_step.py --gemm-mode tf32x3
    python tools/sm_time.py gpurun_out/smtime.csv

With many task slots in flight the throughput of the whole job is bounded by the SUM of SM-time over kernels
(148 SMs x wall time), not by the serialized durations, so this is the ranking that matters for tasks/s."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    col = {n: j for j, n in enumerate(rows[hdr])}
    per = defaultdict(lambda: defaultdict(float))
    ids = defaultdict(set)
    for r in rows[hdr + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        name = re.sub(r"^(void )?mliis::", "", name)
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        m = r[col["Metric Name"]]
        unit = r[col["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if unit.startswith("ns") else (v * 1e3 if unit.startswith("ms") else v)   # -> us
        per[name][m] += v
        ids[name].add(r[col["ID"]])
    tot_sm = sum(d["sm__cycles_active.sum"] for d in per.values())
    tot_us = sum(d["gpu__time_duration.sum"] for d in per.values())
    print("total: %d launches, %.1f us serialized, %.3e SM-cycles active (= %.2f ms of a 148-SM GPU at 1.9 GHz)" % (
        sum(len(v) for v in ids.values()), tot_us, tot_sm, tot_sm / 148 / 1.9e6))
    print("| kernel | launches | dur us | dur %% | SM-cycles active | SM-time %% | avg SMs busy |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["sm__cycles_active.sum"]):
        us, sm = d["gpu__time_duration.sum"], d["sm__cycles_active.sum"]
        busy = sm / (us * 1.9e3) if us > 0 else 0.0        # SM-cycles / elapsed cycles at ~1.9 GHz
        print("| `%s` | %d | %.0f | %.1f | %.3e | %.1f | %.0f |" % (name[:60], len(ids[name]), us, 100 * us / tot_us, sm,
                                                                 100 * sm / tot_sm, busy))


if __name__ == "__main__":
    main(sys.argv[1])
