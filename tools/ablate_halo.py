"""Role ablation of the halo-tile kernels at the bench shapes (timing experiments only: the results are garbage when a role is off).
k_conv_halo: RD_B200_HALO_DEBUG  1 = no halo TMA loads, 2 = no MMAs, 4 = no epilogue stores.
k_wgrad_halo: RD_B200_WGH_DEBUG  1 = no TMA loads, 2 = no MMAs, 4 = no bias sums, 8 = dY box only, 16 = X box only.
  python tools/ablate_halo.py [--batch 16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rd_b200.kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--only", default="")
a = ap.parse_args()
B = a.batch
LAYERS = [("sp6 si", 16 * B, 160, 192, 16, 32), ("sp6 gamma|beta", 16 * B, 160, 192, 32, 64), ("sp6 out", 16 * B, 160, 192, 32, 16),
          ("sp5 gamma|beta", 16 * B, 80, 96, 64, 128), ("sp6 out 7ch", 16 * B, 160, 192, 32, 7), ("ana logits 4ch", 4 * B, 160, 192, 64, 4)]
if a.only:
    LAYERS = [l for l in LAYERS if a.only in l[0]]


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


for name, n, h, w, cin, cout in LAYERS:
    G = 16 if n % 16 == 0 and n >= 256 else 4
    d = K.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, G, 1, 0, 0.2, 0)
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    wt = (torch.randn(G, cout, 9, cin, device="cuda") * 0.05).bfloat16()
    wtT = wt.permute(0, 3, 2, 1).contiguous()
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    dy = torch.randn_like(y)
    dx = torch.empty_like(x)
    dK = torch.empty(G, cout, 9, cin, device="cuda")
    db = torch.zeros(cout, device="cuda")
    line = "%-15s fwd  " % name
    for flag in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ["RD_B200_HALO_DEBUG"] = str(flag)
        line += " dbg%d %.3f" % (flag, timeit(lambda: K.conv2d_fwd(d, x, wt, None, y)))
    print(line)
    if cout % 16:          # narrow outputs: forward only (ops pads dY / the transposed weights for the backward)
        continue
    line = "%-15s dgrad" % name
    for flag in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ["RD_B200_HALO_DEBUG"] = str(flag)
        line += " dbg%d %.3f" % (flag, timeit(lambda: K.conv2d_dgrad(d, dy, wtT, dx)))
    print(line)
    os.environ["RD_B200_HALO_DEBUG"] = "0"
    line = "%-15s wgrad" % name
    for flag in (0, 4, 2, 6, 1, 8, 16, 3, 7, 10, 18):
        os.environ["RD_B200_WGH_DEBUG"] = str(flag)
        line += " dbg%d %.3f" % (flag, timeit(lambda: K.conv2d_wgrad(d, x, dy, dK, db)))
    os.environ["RD_B200_WGH_DEBUG"] = "0"
    print(line)
    del x, wt, wtT, y, dy, dx, dK, db
