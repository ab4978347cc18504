#!/usr/bin/env python3
"""Dynamic opcode histogram of one kernel from an ncu report captured with --import-source on.

usage: ncu -i X.ncu-rep --page source --csv --kernel-name k_fast > k.csv ; python tools/ncu_src_hist.py k.csv [--regions N]
Prints warp-level 'Instructions Executed' per opcode (and pipe guess), plus the top stall samples.
"""
import csv
import sys
from collections import Counter

ALU = {"IADD3", "LOP3", "SHF", "PRMT", "VIMNMX", "VIMNMX3", "FMNMX", "ISETP", "SEL", "VIADD", "VABSDIFF4", "VABSDIFF",
       "IABS", "LEA", "MOV", "FSETP", "POPC", "FLO", "BREV", "IMNMX", "CS2R", "P2R", "R2P", "PLOP3", "FSEL", "SGXT", "BMSK",
       "IDP", "IDP4A", "I2FP", "F2FP", "LOP", "IADD"}
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2", "FFMA2", "FADD2", "FMUL2", "DFMA", "DMUL", "DADD"}


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, samples = Counter(), Counter()
    tot = 0
    for r in rows[hi + 1:]:
        if r and r[0] in ("Kernel Name", "Address"):
            break  # a second launch of the same kernel follows: keep the first only
        if len(r) <= iex:
            continue
        toks = r[isrc].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        base = op.split(".")[0]
        if base == "VIMNMX3" or (base == "VIMNMX" and "3" in op):
            base = "VIMNMX3"
        n = int(r[iex] or 0)
        ops[op if "--full" in sys.argv else base] += n
        samples[base] += int(r[ismp] or 0)
        tot += n
    alu = sum(v for k, v in ops.items() if k.split(".")[0] in ALU)
    fma = sum(v for k, v in ops.items() if k.split(".")[0] in FMA)
    print(f"total warp instructions {tot}  alu-pipe {alu} ({alu / tot:.1%})  fma-pipe {fma} ({fma / tot:.1%})")
    for k, v in ops.most_common(40):
        print(f"  {k:14s} {v:12d} {v / tot:6.1%}   samples {samples[k.split('.')[0]]}")


if __name__ == "__main__":
    main()
