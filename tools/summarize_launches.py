"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("total %.2f ms over %d launches (per-launch times are cold-cache and serialised: compare SHARES)" % (T / 1e6, sum(cnt.values())))
for n, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%9.3f ms %5.1f%% %5d  %s" % (v / 1e6, 100 * v / T, cnt[n], n[:100]))
