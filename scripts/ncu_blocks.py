"""Basic-block view of an `ncu --page source --csv` export: runs of SASS instructions with the same executed count.

    python scripts/ncu_blocks.py source.csv [kernel_index] [min_share_percent]
"""
import collections
import csv
import re
import sys


def main(path, which=0, min_share=1.0):
    rows = list(csv.reader(open(path)))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    blk = rows[idx[which]:idx[which + 1]]
    print(blk[0][1][:100])
    hdr = blk[1]
    i_src, i_exe, i_stall = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    out, cur, n, ops, st, start = [], None, 0, collections.Counter(), 0, 0
    for li, r in enumerate(blk[2:]):
        try:
            e, s = int(r[i_exe]), int(r[i_stall] or 0)
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
        op = m.group(2).split(".")[0] if m else "?"
        if e != cur:
            if cur is not None:
                out.append((start, cur, n, ops, st))
            cur, n, ops, st, start = e, 0, collections.Counter(), 0, li
        n += 1
        ops[op] += 1
        st += s
    out.append((start, cur, n, ops, st))
    tot = sum(c * n for _, c, n, _, _ in out)
    tot_st = sum(s for *_, s in out) or 1
    print(f"total warp instructions {tot / 1e6:.1f}M")
    for start, c, n, ops, s in out:
        if c * n > min_share / 100 * tot:
            print(f"line {start:5d} exec {c / 1e6:7.3f}M x {n:4d} = {c * n / 1e6:7.2f}M ({100 * c * n / tot:4.1f}%) stall {100 * s / tot_st:4.1f}% | "
                  + " ".join(f"{k}:{v}" for k, v in ops.most_common(9)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
