"""Opcode histogram (weighted by executions) + exec-count clusters of one kernel in an ncu report.
python tools/ncu_hot.py prof.ncu-rep k_blur"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
data = [(r[ia].strip(), int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(d[1] for d in data); print("total warp instr", tot, "static", len(data))
h = collections.Counter(); hs = collections.Counter()
def opc(s):
    t = s.split(); op = t[0] if not t[0].startswith("@") else t[1]
    return ".".join(op.split(".")[:2]) if op.startswith("VIMNMX") else op.split(".")[0]
for s, e, sm in data: h[opc(s)] += e; hs[opc(s)] += sm
for op, c in h.most_common(22): print(f"{op:14s} {c:12d} {c/tot:.3f}  samples {hs[op]}")
cl = collections.Counter(e for s, e, _ in data)
print("exec-count clusters (count x n_instr):", sorted(cl.items(), key=lambda kv: -kv[0] * kv[1])[:10])
if len(sys.argv) > 3:
    for s, e, sm in data: print(f"{e:10d} {sm:6d}  {s}")
