#!/usr/bin/env python
"""bench.py — train slices/sec of the representation-disentanglement hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one loop body of the reference (src/main_missing.py:165-284): forward (anatomy / modality
encoders, 16 SPADE decodes, cycle re-encoding), the five loss terms, backward, clip_grad_norm_(1.0) and the
Adam(amsgrad) update, on a synthetic BraTS-shaped 4-contrast batch (28 x 160 x 192 per slice, random-init
weights), bf16 activations / tcgen05 convolutions, fp32 master weights and optimizer state.  Per-GPU batch 16
(= the reference's accumulation target `16 // batch_size == 1`, so every step includes the optimizer).
`value` is device-timed (CUDA events) with the inputs resident in HBM; `e2e` times the public Trainer call with
pinned HOST buffers (H2D of the batch and D2H of the loss vector inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_SLICE_M4 = 278.4e9      # fwd 99.49 + bwd 178.89 GFLOP per slice per train step (M=4)
FLOP_PER_SLICE_M2 = 79.9e9       # the same for M=2 (NCANDA)
METRIC = "train slices/sec (device-timed)"
# --workload: the BASELINE.json configs (the default, `brats`, is config 2 = the configuration the metric is quoted on)
WORKLOADS = {
    "brats": {"contrasts": ["T1", "T1c", "T2", "T2_FLAIR"], "dataset": "BraTS", "flop": FLOP_PER_SLICE_M4, "dropoff": False,
              "label": "BraTS 4-contrast (T1/T1ce/T2/FLAIR) disentanglement training step, bf16, per-GPU batch %d"},
    "brats-dropout": {"contrasts": ["T1", "T1c", "T2", "T2_FLAIR"], "dataset": "BraTS", "flop": FLOP_PER_SLICE_M4, "dropoff": True,
                      "label": "BraTS missing-modality training step with random modality dropout (p = 0.2 per slice, src/util.py:538-542), bf16, "
                               "per-GPU batch %d"},
    "ncanda": {"contrasts": ["T1", "T2"], "dataset": "NCANDA", "flop": FLOP_PER_SLICE_M2, "dropoff": False,
               "label": "NCANDA 2-contrast (T1/T2) disentanglement training step at 160x192, bf16, per-GPU batch %d"},
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": p.get("bf16_tflops_sustained", 1404.5), "bf16_burst": p.get("bf16_tflops", 1661.8),
                "hbm": p.get("hbm_gbs", 6504.1), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region: NVML (nvidia-ml-py) when importable — a function call per sample
    instead of an `nvidia-smi` process every 0.2 s next to the timed host loop — else nvidia-smi."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flag = lambda bit: "Active" if (r & bit) else "Not Active"
        # bit values of nvmlClocksEventReason*: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        return [str(sm), str(mx), "0", flag(0x8), flag(0x40), flag(0x20), flag(0x4)]

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1 if self.nvml is not None else 0.2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for k, n in enumerate(names):
            if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows):
                reasons.append(n)
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def run_reference(args):
    """Reference arm: the reference algorithm on the host cores.  /root/reference (pure Python, not installable:
    no setup.py / pyproject) does not exist on the GPU box, so the timed code is the oracle port (oracle/rd_oracle.py,
    pinned bit-exact against the real reference), faithful per-sample CondConv loop included.  Sample: B=2 slices/step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.params import synth_fill_
    from oracle.rd_oracle import RDOracle, clone_state, train_iteration, DEFAULT_CFG
    import rd_b200.data as rd_data
    from tests.helpers import template_state
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    cfg = dict(DEFAULT_CFG)
    state = clone_state(synth_fill_(template_state(4), seed=1234))
    orc = RDOracle(state, cfg, training=True, batched_condconv=False)
    B = 2
    batch = rd_data.synthetic_batch(B, 4, seed=10)
    eps = rd_data.synthetic_eps(B, 4, 16, seed=11)
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    for _ in range(warm):
        train_iteration(orc, batch, eps, (0, 2))
    t0 = time.perf_counter()
    for _ in range(steps):
        train_iteration(orc, batch, eps, (0, 2))
    dt = (time.perf_counter() - t0) / steps
    v = B / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "slices/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BraTS 4-contrast disentanglement training step (reference CPU path, oracle port)",
                       "batch_per_step": B, "H": 160, "W": 192, "modalities": 4},
            "cpu_baseline": {"value": v, "unit": "slices/s", "cores": cores, "kind": "port",
                             "sample": "%d steps of B=%d slices (fwd+bwd+clip), torch CPU fp32, %d threads" % (steps, B, cores)},
            "e2e": {"value": v, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _measured_traffic(n_images):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r01_dominant_kernel_traffic.json), scaled per image; None when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r01_dominant_kernel_traffic.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        t = json.load(f)
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["images"] * n_images


def time_dominant_kernel(torch, K, B, peaks):
    """Roofline of the dominant kernel (largest share of the step in the ncu launch list): k_conv_halo on the
    full-resolution SPADE gamma|beta convolution (32 -> 64 channels, 3x3, 160x192, 16*B images in 16 weight groups),
    called through rd_conv2d_fwd and timed with CUDA events on the launching stream."""
    from rd_b200.lib import RD_ALGO_TCGEN05
    n, h, w, cin, cout = 16 * B, 160, 192, 32, 64
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    wt = (torch.randn(16, cout, 9, cin, device="cuda") * 0.05).bfloat16()
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    d = K.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, 16, 1, 0, 0.2, RD_ALGO_TCGEN05)
    bias = torch.zeros(cout, device="cuda")          # the layer has a bias (and no activation): time the epilogue the model runs
    for _ in range(3):
        K.conv2d_fwd(d, x, wt, bias, y)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        K.conv2d_fwd(d, x, wt, bias, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * n * h * w * cout * cin * 9
    ach = flops / (ms * 1e-3) / 1e12
    bytes_alg = (x.numel() + y.numel() + wt.numel()) * 2
    # The same layer as the training step runs it: the SPADE modulation IN(z) * (1 + gamma) + beta fused into the epilogue (k_conv_halo<2, 1>:
    # reads a and z, writes gamma and mix).  Reported beside the plain kernel; a failure here must never cost the bench line.
    fused = None
    try:
        C = cout // 2
        z = torch.randn(n, h, w, C, device="cuda").bfloat16()
        mean = torch.zeros(n * C, device="cuda")
        invstd = torch.ones(n * C, device="cuda")
        gamma = torch.empty(n, h, w, C, dtype=torch.bfloat16, device="cuda")
        mix = torch.empty_like(gamma)
        from rd_b200.lib import RD_ALGO_AUTO
        df = K.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, 16, 1, 0, 0.2, RD_ALGO_AUTO)
        if K.conv2d_fwd_spade_supported(df, x):
            for _ in range(3):
                K.conv2d_fwd_spade(df, x, wt, bias, z, mean, invstd, gamma, mix)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                K.conv2d_fwd_spade(df, x, wt, bias, z, mean, invstd, gamma, mix)
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1) / reps
            fb = (x.numel() + z.numel() + gamma.numel() + mix.numel() + wt.numel()) * 2
            fused = {"kernel": "k_conv_halo<2, 1> (the same convolution with the SPADE modulation in its epilogue, as the step runs it)",
                     "ms": fms, "tflops": flops / (fms * 1e-3) / 1e12, "frac_of_bf16_burst": flops / (fms * 1e-3) / 1e12 / peaks["bf16_burst"],
                     "algorithmic_bytes": fb, "hbm_gbs": fb / (fms * 1e-3) / 1e9, "frac_of_hbm": fb / (fms * 1e-3) / 1e9 / peaks["hbm"]}
        del z, gamma, mix
    except Exception as e:  # noqa: BLE001
        fused = {"error": repr(e)[:200]}
    return {"kernel": "k_conv_halo (SPADE sp6 gamma|beta 32->64 3x3 @160x192, %d images, 16 weight groups)" % n, "ms": ms, "fused_in_step": fused,
            "flop_per_launch": flops, "tflops": ach, "frac_of_bf16_burst": ach / peaks["bf16_burst"],
            "algorithmic_bytes": bytes_alg, "hbm_gbs": bytes_alg / (ms * 1e-3) / 1e9,
            "frac_of_hbm": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": _measured_traffic(n)}


def run_ours(args):
    import faulthandler
    if os.environ.get("RD_B200_HANG_DUMP"):
        faulthandler.dump_traceback_later(int(os.environ["RD_B200_HANG_DUMP"]), exit=True)
    import torch
    import torch.distributed as dist
    import rd_b200.config as rd_config
    import rd_b200.data as rd_data
    import rd_b200.kernels as K
    import rd_b200.lib as L
    from rd_b200.ddp import GradReducer
    from rd_b200.trainer import Trainer, build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch
    wl = WORKLOADS[args.workload]
    if args.dropoff:
        wl = dict(wl, dropoff=True)
    torch.manual_seed(10)                      # identical random-init weights on every rank
    cfg = rd_config.default_config(precision=args.precision, batch_size=B, contrast_list=wl["contrasts"], dataset_name=wl["dataset"])
    model = build_model(cfg, dev)
    tr = Trainer(model, cfg, B, use_graph=not args.no_graph)
    if world > 1:
        tr.make_reducer(world)
    M = len(wl["contrasts"])
    nbuf = 4
    host = []
    for k in range(nbuf):                      # pinned host batches (different data per rank and per buffer)
        b = rd_data.synthetic_batch(B, M, seed=10 + 1000 * rank + k, dropoff=wl["dropoff"])
        b = {kk: (v.pin_memory() if torch.is_tensor(v) else v) for kk, v in b.items()}
        e = [t.pin_memory() for t in rd_data.synthetic_eps(B, M, cfg["z_size"], seed=500 + 1000 * rank + k)]
        host.append((b, e))
    pairs = [(0, 2), (3, 1), (1, 0), (2, 3)] if M == 4 else [(0, 1)] * 4

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up: eager iterations + graph capture + W replays
    tr.load_batch(host[0][0], host[0][1], pairs[0])
    for _ in range(tr.graph_warmup + 1 + max(args.warmup, 3)):
        tr.train_iteration()
    if wl["dropoff"]:                          # every dropout pattern of the buffers has gone through the captured graph once
        for k in range(nbuf):
            tr.load_batch(host[k][0], host[k][1], pairs[k % len(pairs)])
            tr.train_iteration()
    sync_all()
    if rank == 0 and args.verbose:
        print("warm-up done, losses", tr.losses_host(), file=sys.stderr)

    resident = []
    if wl["dropoff"]:
        resident = [({kk: (v.to(dev) if torch.is_tensor(v) else v) for kk, v in b.items()}, [t.to(dev) for t in e]) for b, e in host]
    # ---- device-timed region: inputs resident in HBM, K steps
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for k in range(args.steps):
        if wl["dropoff"]:                      # a different missing-modality pattern every step: device-side copy of an HBM-resident batch
            tr.load_batch(resident[k % nbuf][0], resident[k % nbuf][1], pairs[k % len(pairs)])
        tr.train_iteration()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())

    # ---- end-to-end region: pinned host batch -> H2D -> step -> D2H of the loss vector, every step
    # (two untimed pipeline steps first: Trainer.prefetch allocates its staging / pinned buffers and copy stream on first use —
    # set-up cost like the graph capture, not per-step work)
    for k in range(2):
        tr.prefetch(host[k][0], host[k][1], pairs[k])
        tr.train_iteration()
    loss_host = [torch.empty(tr.loss_vec.numel(), dtype=tr.loss_vec.dtype).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    sync_all()
    t0 = time.perf_counter()
    # every step's batch crosses PCIe inside the timed region; like a DataLoader with prefetch, the copy of batch k + 1 is issued
    # (Trainer.prefetch: copy stream + staging buffers) before the host blocks on the result of step k
    # ... and the loss vector of EVERY step is copied to pinned host memory and read, one step behind the launches (an asynchronous
    # logger): the host never idles the GPU between two captured iterations
    loss_log = []
    tr.prefetch(host[0][0], host[0][1], pairs[0])
    for k in range(args.steps):
        tr.train_iteration()                   # swaps the staged batch in (device side), runs the captured iteration
        loss_host[k % 2].copy_(tr.loss_vec, non_blocking=True)     # D2H of this step's result, stream-ordered after the step
        loss_evt[k % 2].record()
        if k + 1 < args.steps:
            b, e = host[(k + 1) % nbuf]
            tr.prefetch(b, e, pairs[(k + 1) % len(pairs)])
        if k > 0:
            loss_evt[(k - 1) % 2].synchronize()
            loss_log.append(loss_host[(k - 1) % 2].tolist())
    loss_evt[(args.steps - 1) % 2].synchronize()
    loss_log.append(loss_host[(args.steps - 1) % 2].tolist())
    assert len(loss_log) == args.steps
    sync_all()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)

    losses = tr.losses_host()
    if rank == 0:
        peaks = _peaks()
        value = world * B * args.steps / (ms_dev * 1e-3)
        e2e_v = world * B * args.steps / e2e_s
        h2d = (B * M * 7 * 160 * 192 + B * 160 * 192 * 2 + B * M + M * B * 16) * 4 + 8
        lp = list(tr.launches_per_graph.values())
        per_graph = (max(lp) if lp else None)           # kernels recorded in one captured iteration
        ach = value / world * wl["flop"] / 1e12              # per-GPU algorithmic TFLOP/s
        dom = time_dominant_kernel(torch, K, B, peaks)
        line = {"metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": wl["label"] % B,
                           "per_gpu_batch": B, "global_batch": world * B, "H": 160, "W": 192, "slab": 7, "modalities": M,
                           "cuda_graph": not args.no_graph, "modality_dropout": bool(wl["dropoff"]),
                           "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2; no explicit flush",
                           "e2e_pipeline": "every step: pinned batch -> H2D on a copy stream (Trainer.prefetch, issued while the previous "
                                           "step computes) -> captured iteration -> D2H of the 9 losses into pinned memory, read by the host "
                                           "one step behind the launches",
                           "parallelism": "dp%d" % world},
                # dominant kernel timed alone -> burst peak; the whole step (278.4 GFLOP per slice) -> sustained peak
                "roofline": {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                             "frac": dom["tflops"] / peaks["bf16_burst"], "traffic": dom["traffic"],
                             "traffic_source": "ncu --set full capture of this kernel on this shape (profiles/r01_dominant_kernel_traffic.json), "
                                               "scaled per image; not re-measured in this run",
                             "peak_source": peaks["source"], "dominant_kernel": dom,
                             "whole_step": {"achieved": ach, "peak": peaks["bf16_sustained"], "frac": ach / peaks["bf16_sustained"],
                                            "basis": "%.1f GFLOP/slice (SURVEY §8d, FlopCounter on the reference step) x slices/s per GPU vs sustained bf16 peak" % (wl["flop"] / 1e9)}},
                "e2e": {"value": e2e_v, "unit": "slices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 9 * 4},
                "gpu_launches": (per_graph or 0) * args.steps,
                "launches_per_step": per_graph,
                "clocks": sampler.summary() if sampler else None,
                "losses": {k: round(v, 5) for k, v in losses.items()}}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured graphs hold NCCL work; destroy_process_group() after that blocks indefinitely (observed on B200 /
        # torch 2.11 / NCCL 2.28.9).  Everything is synchronised and printed: leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_sweep(args):
    """BASELINE config 5: ZeroDose-shaped 4-contrast inference sweep over all 15 non-empty missing-modality subsets (rd_b200.inference).
    A step = one sweep of a resident batch of B slices: 32 (contrast, subset) anatomy encodings and 32 output-decoder rows per slice.
    value = slices/s (each slice evaluated under all 15 subsets); config carries the output rows/s."""
    import torch
    import torch.distributed as dist
    import rd_b200.config as rd_config
    import rd_b200.data as rd_data
    from rd_b200.inference import SweepRunner
    from rd_b200.trainer import build_model
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))      # replicas only: used for the barrier / max time
    dev = torch.device("cuda", local)
    B = args.batch
    torch.manual_seed(10)
    cfg = rd_config.default_config(precision=args.precision, batch_size=B, dataset_name="ZeroDose",
                                   contrast_list=["T1", "T1c", "T2_FLAIR", "ASL"])
    model = build_model(cfg, dev)
    sw = SweepRunner(model, B, use_graph=not args.no_graph, dedup=args.sweep_dedup)
    host = []
    for k in range(2):
        b = rd_data.synthetic_batch(B, 4, seed=10 + 1000 * rank + k)
        host.append({kk: (v.pin_memory() if torch.is_tensor(v) else v) for kk, v in b.items()})
    sw.load(host[0]["inputs"], host[0]["mask_img"])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(3 + max(args.warmup, 3)):
        sw.sweep()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sw.sweep()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())
    # end to end: pinned host batch -> H2D -> sweep -> D2H of every subset's output rows, every step
    out_host = torch.empty(tuple(sw.out.shape), dtype=sw.out.dtype).pin_memory()
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sw.load(host[k % 2]["inputs"], host[k % 2]["mask_img"])
        out_host.copy_(sw.sweep(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    sync_all()
    te = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)
    if rank == 0:
        peaks = _peaks()
        value = world * B * args.steps / (ms_dev * 1e-3)
        flop = len(sw.compute_blocks) * (2.48e9 + 9.96e9)      # 2.48 GFLOP per (slice, contrast) encode + 9.96 per decoded row, rows actually computed
        ach = value / world * flop / 1e12
        d2h = sw.out.numel() * sw.out.element_size()
        line = {"metric": "inference slices/sec over all 15 missing-modality subsets (device-timed)", "value": value, "unit": "slices/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": "ZeroDose 4-contrast MRI -> target synthesis inference sweep over all 15 non-empty missing-modality "
                                       "subsets, eval mode, bf16, per-GPU batch %d (replicas only across GPUs)" % B,
                           "per_gpu_batch": B, "subsets": len(sw.subsets), "rows_per_slice": sw.rows_per_slice,
                           "rows_computed_per_slice": len(sw.compute_blocks),
                           "schedule": ("dedup: only the distinct (contrast, contrast-0-present) rows are computed, the rest gathered" if sw.dedup else
                                        "every (subset, present contrast) pair computed, all 15 subsets batched into one pass"),
                           "output_rows_per_s": value * sw.rows_per_slice, "cuda_graph": not args.no_graph,
                           "l2": "per-sweep working set exceeds the 126 MB L2; no explicit flush", "parallelism": "replicas x%d" % world},
                "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                             "frac": ach / peaks["bf16_sustained"], "traffic": None, "peak_source": peaks["source"],
                             "basis": "%d rows computed per slice x (2.48 + 9.96) GFLOP (SURVEY §6) vs sustained bf16 peak" % len(sw.compute_blocks)},
                "e2e": {"value": world * B * args.steps / e2e_s, "unit": "slices/s",
                        "h2d_bytes_per_step": (B * 28 * 160 * 192 + B * 160 * 192) * 4, "d2h_bytes_per_step": d2h},
                "gpu_launches": (sw.launches or 0) * args.steps, "launches_per_step": sw.launches,
                "clocks": sampler.summary() if sampler else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def cpu_baseline():
    """The oracle port on the box's host cores, bounded sample: 1 warm + 1 timed step of B=2 slices."""
    import torch
    from oracle.params import synth_fill_
    from oracle.rd_oracle import RDOracle, clone_state, train_iteration, DEFAULT_CFG
    import rd_b200.data as rd_data
    from tests.helpers import template_state
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    orc = RDOracle(clone_state(synth_fill_(template_state(4), seed=1234)), dict(DEFAULT_CFG), training=True)
    B = 2
    batch = rd_data.synthetic_batch(B, 4, seed=10)
    eps = rd_data.synthetic_eps(B, 4, 16, seed=11)
    train_iteration(orc, batch, eps, (0, 2))
    t0 = time.perf_counter()
    train_iteration(orc, batch, eps, (0, 2))
    dt = time.perf_counter() - t0
    return {"value": B / dt, "unit": "slices/s", "cores": cores, "kind": "port",
            "sample": "1 warm + 1 timed step of B=2 slices (fwd+bwd+clip), torch CPU fp32, %d threads" % cores}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dropoff", action="store_true", help="random modality dropout (config 3; same as --workload brats-dropout)")
    ap.add_argument("--workload", default="brats", choices=sorted(WORKLOADS) + ["infer-sweep"],
                    help="BASELINE.json configs: brats (2, default), brats-dropout (3), ncanda (4), infer-sweep (5)")
    ap.add_argument("--sweep-dedup", action="store_true", help="infer-sweep: compute only the distinct rows (see rd_b200.inference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "infer-sweep":
        run_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
