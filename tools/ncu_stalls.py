"""Top stall-sampled SASS instructions of an ncu report (source page), with the share of all samples.
  python tools/ncu_stalls.py report.ncu-rep [min_pct] [context]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h, data = rows[k0], rows[k0 + 1:]
iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
shown = set()
for k, r in enumerate(data):
    s = int(r[iS])
    if s > tot * minpct / 100:
        for j in range(max(0, k - ctx), min(len(data), k + ctx + 1)):
            if j in shown:
                continue
            shown.add(j)
            rr = data[j]
            top = sorted(((int(rr[i] or 0), n) for i, n in stall_cols), reverse=True)[:2]
            print("%5d %6d %5.1f%% ex %9s  %-70s %s" % (j, int(rr[iS]), 100 * int(rr[iS]) / tot, rr[iEx], rr[iSrc].strip()[:70],
                                                      " ".join("%s=%d" % (n[6:], v) for v, n in top if v)))
