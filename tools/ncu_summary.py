"""Summarise an ncu launch list (csv) and/or a full report (.ncu-rep):  python tools/ncu_summary.py launches.csv prof.ncu-rep"""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault((row["Kernel Name"].split("(")[0], row["Grid Size"]), []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | grid | launches | mean us | total us | share |\n|---|---|---|---|---|---|")
    for k, v in agg.items():
        print(f"| {k[0]} | {k[1]} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
    print(f"\nTotal {tot/1e3:.1f} us\n")

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp inst"), ("lts__t_sector_hit_rate.pct", "L2 hit%"), ("l1tex__t_sector_hit_rate.pct", "L1 hit%")]

def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = [w for w in WANT if w[0] in hdr]
    print("| kernel | grid | " + " | ".join(w[1] for w in want) + " |\n|---|---|" + "---|" * len(want))
    for r in rows[2:]:
        vals = []
        for w, _ in want:
            v = r[hdr.index(w)]
            try: vals.append(f"{float(v.replace(',', '')):.1f}")
            except ValueError: vals.append(v)
        print(f"| {r[hdr.index('Kernel Name')].split('(')[0]} | {r[hdr.index('Grid Size')]} | " + " | ".join(vals) + " |")

for a in sys.argv[1:]:
    (full if a.endswith(".ncu-rep") else launches)(a)
