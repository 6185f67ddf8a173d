"""Per-reason stall totals and the hottest instructions of each reason from an `ncu --page source --csv` export.

    python scripts/ncu_stalls.py source.csv [kernel_index] [top_n]
"""
import csv
import sys


def main(path, which=0, top=6):
    rows = list(csv.reader(open(path)))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    blk = rows[idx[which]:idx[which + 1]]
    hdr = blk[1]
    print(blk[0][1][:100])
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    i_src, i_exe = hdr.index("Source"), hdr.index("Instructions Executed")
    tot = {}
    lines = {r: [] for r in reasons}
    for li, r in enumerate(blk[2:]):
        for name in reasons:
            try:
                v = int(r[hdr.index(name)] or 0)
            except ValueError:
                v = 0
            if v:
                tot[name] = tot.get(name, 0) + v
                lines[name].append((v, li, r[i_exe], r[i_src][:80]))
    total = sum(tot.values())
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{name:24s} {100 * v / total:5.1f}%")
        if 100 * v / total >= 5:
            for s, li, e, src in sorted(lines[name], reverse=True)[:top]:
                print(f"      {100 * s / total:4.1f}%  line {li:5d} exec {e:>9s}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 6)
