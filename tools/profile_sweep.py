"""One profiled inference sweep (eager launches) after warm-up, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/sweep_launches.csv python tools/profile_sweep.py --batch 16"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rd_b200.config as rd_config
import rd_b200.data as rd_data
from rd_b200.inference import SweepRunner
from rd_b200.trainer import build_model

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--dedup", action="store_true")
a = ap.parse_args()
torch.manual_seed(10)
cfg = rd_config.default_config(precision="bf16", batch_size=a.batch, dataset_name="ZeroDose", contrast_list=["T1", "T1c", "T2_FLAIR", "ASL"])
model = build_model(cfg, "cuda:0")
sw = SweepRunner(model, a.batch, use_graph=False, dedup=a.dedup)
b = rd_data.synthetic_batch(a.batch, 4, seed=10)
sw.load(b["inputs"], b["mask_img"])
for _ in range(2):
    sw.sweep()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
sw.sweep()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one eager sweep: %.2f ms, B=%d, rows computed per slice %d" % (e0.elapsed_time(e1), a.batch, len(sw.compute_blocks)))
