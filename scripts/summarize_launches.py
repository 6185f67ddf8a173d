"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

    python scripts/summarize_launches.py gpurun_out/launches.csv "title" [anchor-regex] > profiles/<name>.md

With an anchor regex (a kernel launched exactly once per step, e.g. "tail.*fwd") only the launches between its last
two occurrences are summarised: exactly one training step of a capture that spans several.
"""
import collections
import csv
import re
import sys


def main(path, title, anchor=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if anchor:
        hits = [i for i, r in enumerate(rows) if re.search(anchor, r["Kernel Name"])]
        if len(hits) >= 2:
            rows = rows[hits[-2] + 1:hits[-1] + 1]
    agg, total = collections.OrderedDict(), 0.0
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    ours = sum(v for k, (n, v) in agg.items() if "pit::" in k)
    print(f"# {title}\n")
    print(f"Source: `{path}` (ncu `--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised launches: compare shares, not absolutes).\n")
    print(f"Total device time in window: {total/1e3:.3f} ms; libpit_posatt.so kernels: {ours/1e3:.3f} ms ({100*ours/total:.1f} %).\n")
    print("| device time (us) | launches | share | kernel |\n|---:|---:|---:|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"| {v:.1f} | {n} | {100*v/total:.1f} % | `{k}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list", sys.argv[3] if len(sys.argv) > 3 else None)
