"""Per-launch table from an `ncu --metrics gpu__time_duration.sum[,...] --csv` log: id, kernel, us, extra metrics."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per, order = {}, []
for row in csv.DictReader(lines):
    k = (int(row["ID"]), row["Kernel Name"])
    per.setdefault(k, {})[row["Metric Name"]] = row["Metric Value"]
    if k not in order:
        order.append(k)
for k in order:
    m = per[k]
    name = re.sub(r"^void |adamvs::", "", re.sub(r"\(.*", "", k[1]))[:72]
    us = float(m.pop("gpu__time_duration.sum", "0").replace(",", "")) / 1e3
    rest = " ".join(f"{v}" for _, v in sorted(m.items()))
    print(f"{k[0]:5d} {name:72s} {us:9.1f} us  {rest}")
