#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON + text block for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x  [--traffic-json profiles/distill_kernel_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_json = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for k in KEYS:
            if k in h:
                i = h.index(k)
                try:
                    d[k] = [float(r[i].replace(",", "")), units[i]]
                except ValueError:
                    d[k] = [r[i], units[i]]
        kernels.append(d)
    json.dump({"report": rep, "kernels": kernels}, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for d in kernels:
            f.write("kernel: %s\n" % d["kernel"])
            for k in KEYS:
                if k in d:
                    f.write("  %-70s %s %s\n" % (k, d[k][0], d[k][1]))
            f.write("\n")
    if traffic_json and kernels:
        def mb(d, k):
            v, u = d[k]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        t = [mb(d, "dram__bytes_read.sum") + mb(d, "dram__bytes_write.sum") for d in kernels]
        json.dump({"dram_bytes_per_launch": sum(t) / len(t), "launches": len(t), "source": rep,
                   "kernel": kernels[0]["kernel"]}, open(traffic_json, "w"), indent=1)
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
