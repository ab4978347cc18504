#!/usr/bin/env python3
"""Per-source-line dynamic instruction counts of one kernel from an ncu report (--import-source on).

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name k_fast > k.csv
       python tools/ncu_line_hist.py k.csv [min_share]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hi]
    iex, ismp = h.index("Instructions Executed"), h.index("# Samples")
    lines = []
    seen_fn = 0
    for r in rows[hi + 1:]:
        if r and r[0] in ("File Path", "Function Name", "Line No"):
            if r[0] == "Function Name":
                seen_fn += 1
                if seen_fn > 0 and lines and "--all" not in sys.argv:
                    break
            continue
        if len(r) <= iex or not r[0]:
            continue
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[iex] or 0), int(r[ismp] or 0)))
        except ValueError:
            pass
    tot = sum(x[2] for x in lines) or 1
    tsm = sum(x[3] for x in lines) or 1
    print(f"total warp instructions {tot}, samples {tsm}")
    for ln, src, n, s in sorted(lines):
        if n / tot >= min_share or s / tsm >= min_share:
            print(f"{ln:5d} {n / tot:6.1%} inst {s / tsm:6.1%} smp  {src[:110]}")


if __name__ == "__main__":
    main()
