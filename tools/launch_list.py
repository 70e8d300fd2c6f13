#!/usr/bin/env python
"""Summarise an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv ...``) per kernel:
launches, total and average duration, share of the listed time.  Per-launch times under ncu are cold-cache and
serialised, so only the SHARES are comparable with an un-profiled run.

    python tools/launch_list.py gpurun_out/launches.csv > profiles/rN_launches.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]          # ncu banner lines
    rows = list(csv.reader(lines))
    hdr = next((r for r in rows if "Kernel Name" in r and "Metric Value" in r), None)
    if hdr is None:
        raise SystemExit("not an ncu long-format CSV (no 'Kernel Name' / 'Metric Value' header)")
    k, name, unit, val = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6, "s": 1e6}
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= val or r[name] != "gpu__time_duration.sum":
            continue
        kernel = re.sub(r"\(.*$", "", r[k]).strip()                  # drop the argument list, keep template arguments
        a = agg[kernel]
        a[0] += 1
        a[1] += float(r[val].replace(",", "")) * scale.get(r[unit], 1.0)
    total = sum(a[1] for a in agg.values()) or 1.0
    print("# launches     total us    avg us   share  kernel")
    for kernel, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:10d} {us:12.1f} {us / n:9.1f} {100 * us / total:6.2f}%  {kernel}")
    print(f"# total {sum(a[0] for a in agg.values())} launches, {total / 1e3:.1f} ms")


if __name__ == "__main__":
    main(sys.argv[1])
