"""k_wgrad_tma split selection against a measured launch list — runs without a GPU.

    python tools/wgrad_split_table.py profiles/r02_launches_step_b16_v37.csv gpurun_out/r02_trace_conv_v37.txt

Joins the ncu launch list of one training iteration with the RD_B200_TRACE_CONV=1 trace of the same program (like tools/conv_table.py)
and prints, per k_wgrad_tma launch, the measured time under the old "two waves" rule, the CTA count of that rule and the split
`rd_wgrad_tma_plan` (the host helper of the ABI: the same code the launch runs) chooses now.  Without arguments it prints the plan of the
step's layer shapes at B = 16."""
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rd_b200.kernels as K
import rd_b200.lib as L


def plan_line(n, g, H, W, cin, cout, k, s, sm=148):
    p = K.wgrad_tma_plan(K.conv_desc(n, H, W, cin, cout, k, k, s, 1, g, L.RD_BF16), sm)
    if p is None:
        return "not on k_wgrad_tma"
    return "two waves: %4d CTAs = %.2f waves | chosen: %2d of %2d X boxes per CTA x %2d chunks = %4d CTAs = %.2f waves" % (
        p["ctas_two_waves"], p["ctas_two_waves"] / sm, p["xb_per_cta"], p["xb_total"], p["chunks_per_group"], p["ctas"], p["ctas"] / sm)


def main():
    if len(sys.argv) < 3:
        for (H, W, cin, cout, k, s, n, g) in [(40, 48, 128, 256, 3, 1, 256, 16), (40, 48, 128, 64, 3, 1, 256, 16), (20, 24, 128, 256, 3, 1, 256, 16),
                                              (20, 24, 128, 128, 3, 1, 256, 16), (40, 48, 256, 64, 3, 1, 64, 4), (20, 24, 512, 128, 3, 1, 64, 4),
                                              (10, 12, 256, 256, 3, 1, 64, 4)]:
            print("n=%d g=%d %dx%d %d->%d k%d s%d  %s" % (n, g, H, W, cin, cout, k, s, plan_line(n, g, H, W, cin, cout, k, s)))
        return
    lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
    t = {}
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum" and "k_wgrad_tma" in r["Kernel Name"]:
            t[int(r["ID"])] = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    times = [t[i] for i in sorted(t)]
    tr = [l.split() for l in open(sys.argv[2]) if l.startswith("rd_conv wgrad tma")][-len(times):]
    assert len(tr) == len(times), (len(tr), len(times))
    for us, l in zip(times, tr):
        kv = dict(x.split("=") for x in l[3:] if "=" in x)
        H, W = map(int, [x for x in l[3:] if re.fullmatch(r"\d+x\d+", x)][0].split("x"))
        print("%7.1f us  n=%-3s g=%-2s %3dx%-3d %3s->%-3s k%s s%s  %s" % (us, kv["n"], kv["g"], H, W, kv["cin"], kv["cout"], kv["k"], kv["s"],
              plan_line(int(kv["n"]), int(kv["g"]), H, W, int(kv["cin"]), int(kv["cout"]), int(kv["k"]), int(kv["s"]))))
    print("total %.1f us under the two-waves rule (ncu per-launch times: cold cache, serialised)" % sum(times))


if __name__ == "__main__":
    main()
