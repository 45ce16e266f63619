"""Join an ncu launch list (tools/profile_step.py under ncu, CSV) with the RD_B200_TRACE_CONV=1 trace of the same program: one row per
convolution launch of the last iteration — kernel, operation, shape, time, algorithmic TFLOP/s, DRAM MB and, when the launch list was taken
with `sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed`, the tensor pipe's share of the launch's cycles.
    python tools/conv_table.py launches.csv trace.txt [iterations_in_trace=3] [top=200]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["t"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    elif r["Metric Name"].startswith("sm__pipe_tensor_cycles_active"):
        d["tp"] = v
    elif r["Metric Name"].startswith("dram__bytes"):
        d["b"] = d.get("b", 0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
convs = [d for d in per.values() if re.search(r"k_conv_halo|k_conv_tma|k_wgrad_halo|k_wgrad_tma|k_conv_tc|k_wgrad_tc|k_conv_sw", d["name"])]
tr = [l.split() for l in open(sys.argv[2]) if l.startswith("rd_conv")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = len(tr) // iters
tr = tr[-n:]
assert len(convs) == n, (len(convs), n)
rows = []
for d, t in zip(convs, tr):
    kv = dict(x.split("=") for x in t[3:] if "=" in x)
    size = [x for x in t[3:] if "=" not in x][0]
    h, w = map(int, size.split("x"))
    N, cin, cout, k, s = int(kv["n"]), int(kv["cin"]), int(kv["cout"]), int(kv["k"]), int(kv["s"])
    fl = 2 * N * (h // s) * (w // s) * cin * cout * k * k
    rows.append((d["t"], t[1], t[2], d["name"].split("::")[-1], N, size, cin, cout, k, s, fl / d["t"] / 1e6, d.get("b", 0) / 1e6, d.get("tp")))
tot = sum(r[0] for r in rows)
print("convolution launches: %d, %.2f ms (ncu per-launch times: cold cache, serialised)" % (n, tot / 1e3))
top = int(sys.argv[4]) if len(sys.argv) > 4 else 200
if rows and rows[0][-1] is not None:
    fw = sum(r[0] * r[-1] for r in rows) / tot
    print("tensor pipe active (sm__pipe_tensor_cycles_active, %% of elapsed cycles), time-weighted over the convolution launches: %.1f %%" % fw)
for r in sorted(rows, key=lambda r: -r[0])[:top]:
    print("%7.1f us %-6s %-10s %-16s n=%d %s cin=%d cout=%d k=%d s=%d  %7.1f TF/s %7.1f MB" % r[:-1] +
          ("" if r[-1] is None else "  tensor pipe %4.1f %%" % r[-1]))
