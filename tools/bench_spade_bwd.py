"""SPADE modulation backward at the bench shapes: single-pass kernel (k_spade_bwd_fused) against the two-pass form
(RD_B200_SPADE_BWD_FUSED=0), CUDA events, tensors far larger than the L2.
  python tools/bench_spade_bwd.py [--batch 16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rd_b200.kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
dev = torch.device("cuda:0")
N = 16 * a.batch
SHAPES = [("sp6", 160, 192, 32), ("sp5", 80, 96, 64), ("sp4", 40, 48, 128), ("sp3", 20, 24, 128), ("sp2", 10, 12, 128), ("sp1", 5, 6, 128)]
tot = {"0": 0.0, "1": 0.0}
for name, h, w, C in SHAPES:
    g = torch.Generator(device=dev).manual_seed(1)
    z = torch.randn(N, h, w, C, device=dev, generator=g).bfloat16()
    gamma = (torch.randn(N, h, w, C, device=dev, generator=g) * 0.5).bfloat16()
    dmix = torch.randn(N, h, w, C, device=dev, generator=g).bfloat16()
    mean, invstd = torch.empty(N * C, device=dev), torch.empty(N * C, device=dev)
    K.norm_stats(z, N, h * w, C, 1e-5, K.norm_workspace(N, h * w, C, dev), mean, invstd, None, None, None, 0.0)
    dz, dgb = torch.empty_like(z), torch.empty(N, h, w, 2 * C, dtype=torch.bfloat16, device=dev)
    ws = K.spade_bwd_workspace(z)
    unit = z.numel() * 2
    line = "%-4s N=%d %dx%d C=%d" % (name, N, h, w, C)
    for mode in ("0", "1"):
        os.environ["RD_B200_SPADE_BWD_FUSED"] = mode
        for _ in range(3):
            K.spade_modulate_bwd_g(z, mean, invstd, gamma, dmix, dz, dgb, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            K.spade_modulate_bwd_g(z, mean, invstd, gamma, dmix, dz, dgb, ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tot[mode] += ms
        units = 9 if mode == "0" else 6
        line += "  | %s %.3f ms, %.2f TB/s over %d tensor units" % ("two-pass" if mode == "0" else "one-pass", ms, units * unit / ms / 1e9, units)
    print(line)
print("sum over the six SPADE blocks: two-pass %.3f ms, one-pass %.3f ms" % (tot["0"], tot["1"]))
