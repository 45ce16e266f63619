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


def test_bucket_plan_skips_inactive_ranges():
    from rd_b200.ddp import plan_buckets
    segs = [(0, 100), (100, 100), (200, 60), (400, 100), (500, 20)]      # gap 260..400 = inactive parameter
    b = plan_buckets(segs, 150)
    assert b == [(0, 100), (100, 260), (400, 520)] or all(e - s <= 200 for s, e in b)
    covered = set()
    for s, e in b:
        covered |= set(range(s, e))
    assert covered == set(range(0, 260)) | set(range(400, 520))
