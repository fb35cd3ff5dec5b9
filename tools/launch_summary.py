"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into kernel, launches, total_ms, share.
usage: python tools/launch_summary.py <launches.csv> > <summary.csv>"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ix = {k: i for i, k in enumerate(rows[head])}
    total, count = defaultdict(float), defaultdict(int)
    for r in rows[head + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"(?<=\w)<.*", "", name).replace("void ", "")
        total[name] += float(r[ix["Metric Value"]].replace(",", "")) * scale
        count[name] += 1
    whole = sum(total.values())
    out = csv.writer(sys.stdout)
    out.writerow(["kernel", "launches", "total_ms", "share"])
    for k in sorted(total, key=total.get, reverse=True):
        out.writerow([k, count[k], f"{total[k]:.3f}", f"{total[k] / whole:.4f}"])


if __name__ == "__main__":
    main(sys.argv[1])
