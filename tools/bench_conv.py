"""Per-layer timing of the convolution kernels (CUDA events) on the shape families of SURVEY.md §8a at the
bench batch: forward, dgrad, wgrad; prints TFLOP/s and the fraction of the measured bf16 peak.
  python tools/bench_conv.py [--batch 16] [--json out.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rd_b200.kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--json", default="")
ap.add_argument("--algo", type=int, default=0)
ap.add_argument("--only", default="", help="substring of the layer name")
a = ap.parse_args()
B = a.batch
peak = 1661.8
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.isfile(pp):
    peak = json.load(open(pp)).get("bf16_tflops", peak)
# name, images, H, W, Cin, Cout, k, stride, pad, groups
LAYERS = [
    ("sp6 gamma|beta", 16 * B, 160, 192, 32, 64, 3, 1, 1, 16),
    ("sp6 out", 16 * B, 160, 192, 32, 16, 3, 1, 1, 16),
    ("sp6 si", 16 * B, 160, 192, 16, 32, 3, 1, 1, 16),
    ("sp5 gamma|beta", 16 * B, 80, 96, 64, 128, 3, 1, 1, 16),
    ("sp5 out", 16 * B, 80, 96, 64, 32, 3, 1, 1, 16),
    ("sp5 si", 16 * B, 80, 96, 16, 64, 3, 1, 1, 16),
    ("sp4 gamma|beta", 16 * B, 40, 48, 128, 256, 3, 1, 1, 16),
    ("sp4 out", 16 * B, 40, 48, 128, 64, 3, 1, 1, 16),
    ("sp3 gamma|beta", 16 * B, 20, 24, 128, 256, 3, 1, 1, 16),
    ("sp3 out", 16 * B, 20, 24, 128, 128, 3, 1, 1, 16),
    ("sp2 gamma|beta", 16 * B, 10, 12, 128, 256, 3, 1, 1, 16),
    ("sp1 gamma|beta", 16 * B, 5, 6, 128, 256, 3, 1, 1, 16),
    ("ana bottleneck", 4 * B, 10, 12, 256, 256, 3, 1, 1, 4),
    ("ana dec up_4", 4 * B, 20, 24, 512, 128, 3, 1, 1, 4),
    ("ana dec up_1", 4 * B, 160, 192, 128, 32, 3, 1, 1, 4),
    ("ana dec up_2", 4 * B, 80, 96, 256, 64, 3, 1, 1, 4),
    ("ana dec up_3", 4 * B, 40, 48, 512, 128, 3, 1, 1, 4),
    ("ana enc down_2", 4 * B, 80, 96, 32, 64, 4, 2, 1, 4),
    ("ana enc down_4", 4 * B, 20, 24, 128, 256, 4, 2, 1, 4),
]
rows = []


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, n, h, w, cin, cout, k, st, pad, G in LAYERS:
    if a.only and a.only not in name:
        continue
    d = K.conv_desc(n, h, w, cin, cout, k, k, st, pad, G, 1, 0, 0.2, a.algo)
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    wt = (torch.randn(G, cout, k * k, cin, device="cuda") * 0.05).bfloat16()
    wtT = wt.permute(0, 3, 2, 1).contiguous()
    y = torch.empty(n, d.oh, d.ow, cout, dtype=torch.bfloat16, device="cuda")
    dy = torch.randn_like(y)
    dx = torch.empty_like(x)
    dK = torch.empty(G, cout, k * k, cin, device="cuda")
    db = torch.zeros(cout, device="cuda")        # the step always asks for the bias gradient too
    flops = 2.0 * n * d.oh * d.ow * cout * cin * k * k
    t_f = timeit(lambda: K.conv2d_fwd(d, x, wt, None, y))
    if "gamma|beta" in name:       # SPADE block: convolution + modulation pass against the fused epilogue
        Cz = cout // 2
        z = torch.randn(n, h, w, Cz, device="cuda").bfloat16()
        mean, inv = torch.zeros(n * Cz, device="cuda"), torch.ones(n * Cz, device="cuda")
        gam, mix = torch.empty_like(z), torch.empty_like(z)
        bias = torch.zeros(cout, device="cuda")
        t_m = timeit(lambda: K.spade_modulate_fwd(z, mean, inv, y, mix))
        if K.conv2d_fwd_spade_supported(d, x):
            t_s = timeit(lambda: K.conv2d_fwd_spade(d, x, wt, bias, z, mean, inv, gam, mix))
            print("%-16s SPADE: conv %.3f + modulate %.3f = %.3f ms   fused epilogue %.3f ms (%.0f GB/s algorithmic)"
                  % (name, t_f, t_m, t_f + t_m, t_s, (x.numel() + 3 * z.numel()) * 2 / t_s / 1e6))
        del z, gam, mix
    t_d = timeit(lambda: K.conv2d_dgrad(d, dy, wtT, dx))
    t_w = timeit(lambda: K.conv2d_wgrad(d, x, dy, dK, db))
    io = (x.numel() + y.numel()) * 2
    r = {"layer": name, "gflop": flops / 1e9, "fwd_ms": t_f, "dgrad_ms": t_d, "wgrad_ms": t_w,
         "fwd_tflops": flops / t_f / 1e9, "dgrad_tflops": flops / t_d / 1e9, "wgrad_tflops": flops / t_w / 1e9,
         "fwd_frac_peak": flops / t_f / 1e9 / peak, "fwd_io_gbs": io / t_f / 1e6}
    rows.append(r)
    print("%-16s %7.1f GF  fwd %7.3f ms %6.1f TF/s (%4.1f%%, io %5.0f GB/s) | dgrad %7.3f ms %6.1f TF/s | wgrad %7.3f ms %6.1f TF/s"
          % (name, r["gflop"], t_f, r["fwd_tflops"], 100 * r["fwd_frac_peak"], r["fwd_io_gbs"], t_d, r["dgrad_tflops"], t_w, r["wgrad_tflops"]))
    del x, wt, wtT, y, dy, dx, dK, db
tot = sum(r["gflop"] for r in rows)
tt = sum(r["fwd_ms"] + r["dgrad_ms"] + r["wgrad_ms"] for r in rows)
print("FLOP-weighted: %.1f TF/s over fwd+dgrad+wgrad of these layers (%.1f%% of burst bf16 peak %.0f)" % (3 * tot / tt, 300 * tot / tt / peak, peak))
if a.json:
    json.dump(rows, open(a.json, "w"), indent=1)
