"""Data-parallel parity on the REAL model (SURVEY §8e / VERDICT r01 item 7a), run as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_parity.py

Every rank runs the reference step on its own shard of B slices (fp32 parity kernels, the same deterministic weights on all
ranks); the flat gradient buffer is averaged by `GradReducer` (bucketed NCCL all-reduce, early buckets launched from the
tape marker).  Checks, printed by rank 0 as one JSON line:
  (1) averaged gradients == mean over ranks of the CPU ORACLE's shard gradients (oracle/rd_oracle.py, pinned bit-identical
      to the reference), every parameter, 1e-3 of the parameter's gradient scale; same grad-None set;
  (2) clipped-gradient norm == the norm of the oracle's mean gradient;
  (3) after `--steps` full iterations (forward, backward, all-reduce, clip, Adam; bf16 product kernels, CUDA graph with the
      NCCL all-reduces captured inside) parameters, Adam moments and step counters are BIT-EQUAL on all ranks, while the
      BatchNorm running buffers (rank-local by design) differ.
The oracle is the checker only (this is a tool, like tests/)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1, help="slices per rank (the oracle runs world x batch slices on the CPU)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import rd_b200.config as rd_config
    import rd_b200.data as rd_data
    from rd_b200.trainer import Trainer, build_model
    from oracle.params import synth_fill_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, M = args.batch, 4
    pair = (0, 2)
    report = {"world": world, "per_rank_batch": B}

    def shard(r, k=0):
        rows = [[1, 1, 1, 1]] * B if (r + k) % 2 == 0 else [[1, 0, 1, 1]] + [[1, 1, 1, 1]] * (B - 1)
        batch = rd_data.synthetic_batch(B, M, seed=40 + 7 * r + 100 * k, missing=rows, zero_border=8)
        eps = rd_data.synthetic_eps(B, M, 16, seed=41 + 7 * r + 100 * k)
        return batch, eps

    # ------------------------------------------------------------ (1), (2): fp32, eager, one backward + all-reduce
    cfg = rd_config.default_config(precision="fp32", batch_size=B)
    torch.manual_seed(100 + rank)            # deliberately DIFFERENT initial weights per rank: make_reducer must broadcast rank 0's
    model = build_model(cfg, dev)
    if rank == 0:
        state = synth_fill_({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, seed=1234)
        model.load_state_dict(state)
    tr = Trainer(model, cfg, B, use_graph=False)
    tr.make_reducer(world)
    batch, eps = shard(rank)
    tr.load_batch(batch, eps, pair)
    model.train()
    tr._fwd_bwd()
    tr.ddp.finish(tr.fp)
    torch.cuda.synchronize()
    fp = tr.fp
    got = {n: fp.grad[o:o + p.numel()].view_as(p).detach().cpu().clone() for n, p, o in zip(fp.names, fp.params, fp.offsets)}
    tr._clip_step(False)
    gnorm = tr.grad_norm_host()
    if rank == 0:
        from oracle.rd_oracle import RDOracle, clone_state, DEFAULT_CFG, param_keys
        mean = None
        for r in range(world):
            orc = RDOracle(clone_state(state), dict(DEFAULT_CFG), training=True, batched_condconv=True)
            b, e = shard(r)
            out = orc.forward_losses(b["inputs"], b["targets"], b["mask"], b["mask_img"], e, pair)
            out["all"].backward()
            g = {k: orc.P[k].grad for k in param_keys(orc.P)}
            if mean is None:
                mean = {k: (None if v is None else v.clone() / world) for k, v in g.items()}
            else:
                for k, v in g.items():
                    if v is not None:
                        mean[k] = v / world if mean[k] is None else mean[k] + v / world
        worst, bad, none_mismatch, checked = 0.0, [], [], 0
        ref_norm = float(torch.sqrt(sum((v.double() ** 2).sum() for v in mean.values() if v is not None)))
        floor = 1e-7 * max(1.0, ref_norm)     # biases that feed a BatchNorm have an exactly zero gradient: what both sides hold there is summation noise of the (unclipped) gradient scale
        act = dict(zip(fp.names, fp.active_mask))
        for k, ref in mean.items():
            if ref is None:
                if act[k] and float(got[k].abs().max()) != 0.0:
                    none_mismatch.append(k)
                continue
            if not act[k]:
                none_mismatch.append(k)
                continue
            scale = float(ref.abs().max())
            err = float((got[k] - ref).abs().max())
            rel = err / max(scale, 1e-30)
            checked += 1
            if err > 1e-3 * scale + floor:
                bad.append((k, err, scale))
            if scale > 1e-6:
                worst = max(worst, rel)
        report["avg_grad_vs_oracle_shard_mean"] = {"parameters_checked": checked, "worst_rel_err": worst, "tolerance": "1e-3 * max|g| + 1e-7 * max(1, ||g||)",
                                                    "failed": [(k, "%.3e" % e, "%.3e" % s) for k, e, s in bad[:10]],
                                                    "grad_none_mismatch": none_mismatch[:10]}
        report["grad_norm"] = {"device": gnorm, "oracle_mean": ref_norm, "rel_err": abs(gnorm - ref_norm) / ref_norm}
    del tr, model
    torch.cuda.empty_cache()

    # ------------------------------------------------------------ (3): product path (bf16, CUDA graph, NCCL captured), full steps
    cfg = rd_config.default_config(precision="bf16", batch_size=B)
    cfg["lr"] = 1e-3
    torch.manual_seed(200 + rank)            # again different per rank before the broadcast
    model = build_model(cfg, dev)
    tr = Trainer(model, cfg, B, use_graph=True)
    tr.accum_every = 1                       # an optimizer step every iteration (the reference steps every 16 // B)
    tr.graph_warmup = 2
    tr.make_reducer(world)
    n_it = tr.graph_warmup + args.steps
    for k in range(n_it):
        b, e = shard(rank, k)
        tr.load_batch(b, e, pair)
        tr.train_iteration()
    torch.cuda.synchronize()
    fp = tr.fp

    def spread(t):
        hi, lo = t.clone(), t.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return int((hi != lo).sum().item())
    bn = torch.cat([b.reshape(-1).float() for n, b in model.named_buffers() if n.endswith("running_mean")])
    eq = {"params_differing_elements": spread(fp.flat), "adam_m": spread(fp.m), "adam_v": spread(fp.v), "adam_vmax": spread(fp.vmax),
          "param_steps": spread(fp.param_steps), "bn_running_mean_differing_elements (rank-local, expected > 0)": spread(bn),
          "iterations": n_it, "graph_replays": max(0, n_it - tr.graph_warmup - 0), "nccl_in_graph": bool(tr.ddp_in_graph),
          "flat_elements": int(fp.flat.numel()), "loss_all_rank0": tr.losses_host()["all"],
          "readiness_stages": {n: {"expected_accumulations": e, "buckets": len(b)} for n, e, b in
                               zip(tr.ddp.stage_names, tr.ddp._expected, tr.ddp.stage_buckets)},
          "stage_all_reduces_launched_from_markers (eager iterations)": tr.ddp.early_launches,
          "late_bucket_bytes": sum(e - s for s, e in tr.ddp.late) * 4}
    report["after_steps"] = eq
    ok = None
    if rank == 0:
        a = report["avg_grad_vs_oracle_shard_mean"]
        ok = (not a["failed"] and not a["grad_none_mismatch"] and report["grad_norm"]["rel_err"] < 1e-3
              and all(eq[k] == 0 for k in ("params_differing_elements", "adam_m", "adam_v", "adam_vmax", "param_steps")))
        report["ok"] = bool(ok)
        line = json.dumps(report)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if (ok is None or ok) else 1)


if __name__ == "__main__":
    main()
