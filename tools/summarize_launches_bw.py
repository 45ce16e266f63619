"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) by kernel:
time, share, launches, DRAM bytes and achieved DRAM GB/s.   python tools/summarize_launches_bw.py launches.csv [top]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["t"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    elif r["Metric Name"].startswith("dram__bytes"):
        d["b"] = d.get("b", 0.0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += d.get("t", 0.0); a[1] += d.get("b", 0.0); a[2] += 1
T = sum(a[0] for a in agg.values())
print("total %.2f ms over %d launches (per-launch times are cold-cache and serialised: compare SHARES)" % (T / 1e3, len(per)))
for n, a in sorted(agg.items(), key=lambda x: -x[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%8.3f ms %5.1f%% %4d  %8.1f MB %6.0f GB/s  %s" % (a[0] / 1e3, 100 * a[0] / T, a[2], a[1] / 1e6, a[1] / a[0] / 1e3 if a[0] else 0, n[:90]))
