"""CPU, world_size 2 over gloo: the data-parallel path (rd_b200/ddp.py + trainer).  Semantics under test
(SURVEY §8e): every rank runs the reference step on its own shard; gradients are AVERAGED over ranks before
clip + Adam; parameters stay bit-identical on all ranks; inactive parameters are never communicated."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import emul
    emul.install()
    from rd_b200.ddp import GradReducer, plan_buckets, ready_marker
    from rd_b200.trainer import FlatParams
    import rd_b200.kernels as K

    torch.manual_seed(0)                               # same weights on every rank

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.randn(300, 7))
            self.unused = torch.nn.Parameter(torch.randn(50))
            self.b = torch.nn.Parameter(torch.randn(1000))
            self.c = torch.nn.Parameter(torch.randn(3, 3, 3))

    m = Tiny()
    fp = FlatParams(m)
    fp.set_active([True, False, True, True])
    # early range = parameter `b`: its buckets are reduced from the tape marker's backward (overlap with the rest of backward)
    ob = fp.offsets[2]
    red = GradReducer(fp, world, bucket_mb=0.004, early_range=(ob, ob + 1000))
    fired = []
    hyper = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-5, 0.0, 0, 0])
    g_all = []
    for step in range(3):
        g = torch.Generator().manual_seed(100 * step + rank)       # different "shard" per rank
        for p in (m.a, m.b, m.c):
            p.grad.copy_(torch.randn(p.shape, generator=g))
        m.unused.grad.fill_(float(rank + 1))                        # must never be reduced
        g_all.append(fp.grad.clone())
        if step == 1:
            # the marker fires when the backward of everything downstream of it has run; `b`'s gradient is final then
            x = torch.ones(3, requires_grad=True)
            (y,) = ready_marker(red.early_ready, x)
            (y * 2).sum().backward()
            fired.append(red._launched[0] and len(red._works) == len(red.early) > 0)
        if step == 2:
            # accumulation-count guard: iteration 0 taught the reducer that stage 0 sees no ops._sink report; after one report the
            # marker must NOT launch the stage early (its gradient may not be final), finish() reduces it, and the stage is then
            # barred from early launches for good (the tape changed between iterations)
            red.note_sink(m.b)
            fired.append(red.stage_ready(0) is False and len(red._works) == 0)
        red.finish(fp)
        K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
        K.grad_scale(fp.grad, fp.segments, fp.nseg, fp.scalars)
        K.adam_amsgrad(fp.flat, fp.grad, fp.m, fp.v, fp.vmax, fp.segments, fp.nseg, hyper)
    assert fired == [True, True], fired
    q.put((rank, fp.flat.numpy().copy(), [g.numpy().copy() for g in g_all], m.unused.grad.numpy().copy(), red.buckets,
           red.bytes_per_step))
    dist.destroy_process_group()


def test_ddp_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, flat0, g0, un0, buckets, nbytes), (r1, flat1, g1, un1, _, _) = res
    flat0, flat1, un0, un1 = (torch.from_numpy(t) for t in (flat0, flat1, un0, un1))
    g0, g1 = [torch.from_numpy(t) for t in g0], [torch.from_numpy(t) for t in g1]
    assert torch.equal(flat0, flat1), "parameters must be bit-identical on all ranks after averaged updates"
    assert float(un0[0]) == 1.0 and float(un1[0]) == 2.0, "inactive parameters are not communicated"
    assert (300 * 7 + 1000 + 27) * 4 <= nbytes <= (300 * 7 + 1000 + 27 + 8) * 4      # only active ranges (+ alignment padding)
    # single-process restatement: average the two shards' grads, clip, Adam — must give the same parameters
    sys.path.insert(0, ROOT)
    from tests import emul
    saved = emul.install()
    try:
        from rd_b200.trainer import FlatParams
        import rd_b200.kernels as K
        torch.manual_seed(0)

        class Tiny(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a = torch.nn.Parameter(torch.randn(300, 7))
                self.unused = torch.nn.Parameter(torch.randn(50))
                self.b = torch.nn.Parameter(torch.randn(1000))
                self.c = torch.nn.Parameter(torch.randn(3, 3, 3))
        m = Tiny()
        fp = FlatParams(m)
        fp.set_active([True, False, True, True])
        hyper = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-5, 0.0, 0, 0])
        for step in range(3):
            fp.grad.copy_(0.5 * (g0[step] + g1[step]))
            K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
            K.grad_scale(fp.grad, fp.segments, fp.nseg, fp.scalars)
            K.adam_amsgrad(fp.flat, fp.grad, fp.m, fp.v, fp.vmax, fp.segments, fp.nseg, hyper)
        act = torch.ones_like(fp.flat, dtype=torch.bool)
        o = fp.offsets[1]
        act[o:o + 50] = False
        assert torch.allclose(fp.flat[act], flat0[act], rtol=1e-6, atol=1e-7)
    finally:
        emul.uninstall(saved)


def _shard(r, k=0):
    """Rank r's slice of iteration k (different masks and images per rank; the same recipe as tools/ddp_parity.py)."""
    import rd_b200.data as rd_data
    rows = [[1, 1, 1, 1]] if (r + k) % 2 == 0 else [[1, 0, 1, 1]]
    batch = rd_data.synthetic_batch(1, 4, seed=40 + 7 * r + 100 * k, missing=rows, zero_border=8)
    eps = rd_data.synthetic_eps(1, 4, 16, seed=41 + 7 * r + 100 * k)
    return batch, eps


def _real_model_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import emul
    emul.install()
    import rd_b200.config as rd_config
    from rd_b200.trainer import Trainer, build_model
    from oracle.params import synth_fill_

    cfg = rd_config.default_config(precision="fp32", batch_size=1)
    torch.manual_seed(100 + rank)            # deliberately DIFFERENT initial weights per rank: make_reducer must broadcast rank 0's
    model = build_model(cfg, "cpu")
    if rank == 0:
        state = synth_fill_({k: v.detach().clone() for k, v in model.state_dict().items()}, seed=1234)
        model.load_state_dict(state)
    tr = Trainer(model, cfg, 1, use_graph=False)
    tr.accum_every = 1
    tr.make_reducer(world)
    fp = tr.fp
    model.train()
    # iteration 0 by hand (forward, backward, all-reduce) so that the averaged gradients can be read before the clip
    batch, eps = _shard(rank)
    tr.load_batch(batch, eps, (0, 2))
    tr._fwd_bwd()
    tr.ddp.finish(fp)
    got = {n: fp.grad[o:o + p.numel()].view_as(p).detach().clone().numpy() for n, p, o in zip(fp.names, fp.params, fp.offsets)}
    tr._clip_step(True)
    gnorm = tr.grad_norm_host()
    tr.ddp.reset()
    # iteration 1 through the public loop body: the tape markers may now release their stages early (counts learned in iteration 0)
    batch, eps = _shard(rank, 1)
    tr.train_iteration(batch, eps, (1, 3))
    bn = torch.cat([b.reshape(-1).float() for n, b in model.named_buffers() if n.endswith("running_mean")])
    q.put((rank, got, gnorm, fp.flat.numpy().copy(), fp.m.numpy().copy(), fp.param_steps.numpy().copy(), bn.numpy().copy(),
           dict(zip(fp.names, [bool(a) for a in fp.active_mask])), int(tr.ddp.early_launches), len(tr.ddp.stage_names)))
    dist.destroy_process_group()


def test_ddp_real_model_world2_gloo():
    """SURVEY 8e on the REAL model (the CPU twin of tools/ddp_parity.py): two ranks, one slice each, host logic over the emulated kernel
    layer.  (1) the all-reduced gradients equal the mean of the CPU oracle's per-shard gradients (the oracle is pinned bit-identical to the
    reference), same grad-None set; (2) the clipped norm is the norm of that mean; (3) after two optimizer steps the parameters, Adam
    moments and step counters are bit-equal on both ranks although they were seeded differently (rank 0's state is broadcast), while the
    BatchNorm running buffers stay rank-local."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_real_model_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    # the oracle's shard gradients while the ranks work
    from oracle.rd_oracle import RDOracle, clone_state, DEFAULT_CFG, param_keys
    from oracle.params import synth_fill_
    from tests.helpers import template_state
    state = synth_fill_({k: v.clone() for k, v in template_state(4).items()}, seed=1234)
    mean = {}
    for r in range(2):
        orc = RDOracle(clone_state(state), dict(DEFAULT_CFG), training=True, batched_condconv=True)
        b, e = _shard(r)
        out = orc.forward_losses(b["inputs"], b["targets"], b["mask"], b["mask_img"], e, (0, 2))
        out["all"].backward()
        for k in param_keys(orc.P):
            g = orc.P[k].grad
            if g is not None:
                mean[k] = g / 2 if mean.get(k) is None else mean[k] + g / 2
            else:
                mean.setdefault(k, None)
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, got0, gn0, flat0, m0, st0, bn0, active, early0, nstage), (_, got1, gn1, flat1, m1, st1, bn1, _, early1, _) = res
    ref_norm = float(torch.sqrt(sum((v.double() ** 2).sum() for v in mean.values() if v is not None)))
    floor = 1e-7 * max(1.0, ref_norm)          # exactly-zero gradients (biases in front of a normalisation): summation noise on both sides
    checked = 0
    for k, ref in mean.items():
        g = torch.from_numpy(got0[k])
        assert (got0[k] == got1[k]).all(), "the averaged gradient differs between the ranks: " + k
        if ref is None:
            assert (not active[k]) or float(g.abs().max()) == 0.0, "gradient where the reference has None: " + k
            continue
        assert active[k], "the reference has a gradient for the inactive parameter " + k
        scale = float(ref.abs().max())
        err = float((g - ref).abs().max())
        assert err <= 1e-3 * scale + floor, (k, err, scale)
        checked += 1
    assert checked > 300, checked
    assert abs(gn0 - ref_norm) <= 1e-3 * ref_norm and gn0 == gn1, (gn0, gn1, ref_norm)
    assert (flat0 == flat1).all() and (m0 == m1).all() and (st0 == st1).all(), "parameters / Adam state must be bit-equal on all ranks"
    assert st0.max() == 2, "two optimizer steps were taken"
    assert (bn0 != bn1).any(), "BatchNorm running buffers are rank-local by design (different shards)"
    assert early0 == early1 and 0 < early0 <= nstage, "the readiness stages were released from the tape markers in the second iteration"


def test_bucket_plan_skips_inactive_ranges():
    from rd_b200.ddp import plan_buckets
    segs = [(0, 100), (100, 100), (200, 60), (400, 100), (500, 20)]      # gap 260..400 = inactive parameter
    b = plan_buckets(segs, 150)
    assert b == [(0, 100), (100, 260), (400, 520)] or all(e - s <= 200 for s, e in b)
    covered = set()
    for s, e in b:
        covered |= set(range(s, e))
    assert covered == set(range(0, 260)) | set(range(400, 520))
