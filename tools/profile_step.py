"""One profiled training iteration (eager launches, so every kernel is visible to ncu) after warm-up.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --batch 4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rd_b200.config as rd_config
import rd_b200.data as rd_data
from rd_b200.trainer import Trainer, build_model

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
torch.manual_seed(10)
cfg = rd_config.default_config(precision=a.precision, batch_size=a.batch)
model = build_model(cfg, "cuda:0")
tr = Trainer(model, cfg, a.batch, use_graph=False)
tr.accum_every = 1
batch = rd_data.synthetic_batch(a.batch, 4, seed=10)
eps = rd_data.synthetic_eps(a.batch, 4, 16, seed=11)
for _ in range(2):
    tr.train_iteration(batch, eps, (0, 2))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
tr.train_iteration(batch, eps, (0, 2))
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one eager iteration: %.2f ms, B=%d, losses %s" % (e0.elapsed_time(e1), a.batch, tr.losses_host()))
