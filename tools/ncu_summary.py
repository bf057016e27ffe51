#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: per launch duration, DRAM bytes, DRAM / tensor / SM
utilisation, registers, grid.  usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py [name-filter]"""
import csv
import re
import sys

WANT = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "lts__t_bytes.sum": "l2_MB",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
}


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    rows = list(csv.reader(sys.stdin))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    if hdr is None:
        print("no header found")
        return
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: j for j, n in enumerate(names)}
    print("| kernel | grid | regs | dur us | DRAM rd MB | DRAM wr MB | DRAM % | L2 MB | SM % | tensor % | occ % |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("mliis::", "")
        if flt and flt not in k:
            continue

        def get(metric, scale=1.0):
            if metric not in col:
                return float("nan")
            v = r[col[metric]].replace(",", "")
            try:
                x = float(v)
            except ValueError:
                return float("nan")
            u = units[col[metric]]
            if u == "ns":
                x /= 1e3
            elif u == "ms":
                x *= 1e3
            elif u == "Kbyte":
                x /= 1e3
            elif u == "byte":
                x /= 1e6
            elif u == "Gbyte":
                x *= 1e3
            return x * scale
        print("| `%s` | %s | %s | %.1f | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
            k[:48], r[col.get("launch__grid_size", 0)] if "launch__grid_size" in col else "?",
            r[col["launch__registers_per_thread"]] if "launch__registers_per_thread" in col else "?",
            get("gpu__time_duration.sum"), get("dram__bytes_read.sum"), get("dram__bytes_write.sum"),
            get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), get("lts__t_bytes.sum"),
            get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            get("sm__warps_active.avg.pct_of_peak_sustained_active")))


if __name__ == "__main__":
    main()
