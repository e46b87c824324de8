#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name:
    python scripts/launch_summary.py gpurun_out/launches.csv [out.txt] [--native-share]
Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.
--native-share appends the split between this repository's kernels (namespace sad::) and library kernels (cuDNN / CUTLASS / ATen)."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
    h = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        if r is h or len(r) <= vi or r[ki] == "Kernel Name":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = re.sub(r"\(.*", "", r[ki])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    lines = ["%-70s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share")]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-70s %6d %12.1f %10.2f %6.1f%%" % (k[:70], n, t, t / n, 100 * t / total))
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    if "--native-share" in sys.argv:
        nat = [(n, t) for k, (n, t) in agg.items() if "sad::" in k]
        lib = [(n, t) for k, (n, t) in agg.items() if "sad::" not in k]
        lines.append("")
        lines.append("native (sad::)   %5d launches %10.1f us %5.1f%%" % (sum(n for n, _ in nat), sum(t for _, t in nat), 100 * sum(t for _, t in nat) / total))
        lines.append("library kernels  %5d launches %10.1f us %5.1f%%" % (sum(n for n, _ in lib), sum(t for _, t in lib), 100 * sum(t for _, t in lib) / total))
    out = "\n".join(lines)
    print(out)
    if args:
        open(args[0], "w").write(out + "\n")


if __name__ == "__main__":
    main()
