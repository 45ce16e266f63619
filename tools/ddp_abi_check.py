"""rd_ddp_* through the C ABI (NCCL resolved by the library itself), run under torchrun with one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_abi_check.py
torch.distributed (gloo) is used ONLY to hand rank 0's ncclUniqueId to the other ranks — the collective itself is rd_ddp_bucket_allreduce.
Checks: the averaged bucket equals the mean of the ranks' inputs bit for bit on every rank (integer-valued data), rd_ddp_broadcast
delivers rank 0's bytes, the all-reduce is capturable in an rd_graph and replays correctly.  Prints one JSON line from rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import rd_b200.kernels as K

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ok, ver = K.ddp_available(local)
assert ok, "libnccl.so.2 not loadable"
uid = [K.ddp_unique_id(local) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
K.ddp_init(world, rank, uid[0], local)
dev = torch.device("cuda", local)
n = 5_000_003
g = torch.Generator().manual_seed(7)
base = torch.randint(-8, 9, (n,), generator=g).float()
x = (base * (rank + 1)).to(dev)                                   # rank r holds (r + 1) * base: the mean is base * (world + 1) / 2
K.ddp_bucket_allreduce(x[:n - 3], average=True)                    # a bucket = a contiguous range of the flat buffer
torch.cuda.synchronize()
want = base * (world + 1) / 2.0
res = {"world": world, "nccl_version": ver,
       "allreduce_avg_exact": bool(torch.equal(x[:n - 3].cpu(), want[:n - 3])),
       "outside_bucket_untouched": bool(torch.equal(x[n - 3:].cpu(), (base * (rank + 1))[n - 3:]))}
b = torch.full((1000,), float(rank + 5), device=dev)
K.ddp_broadcast(b, 0)
torch.cuda.synchronize()
res["broadcast_ok"] = bool((b == 5.0).all().item())
# the collective inside an ABI-captured CUDA graph, replayed twice
y = torch.full((4096,), float(rank + 1), device=dev)
s = torch.cuda.Stream(device=dev)
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    gph = K.AbiGraph(local)
    with gph.capture():
        K.ddp_bucket_allreduce(y, average=False)
    gph.launch()
    gph.launch()
    s.synchronize()
tot = world * (world + 1) / 2.0
res["graph_allreduce_ok"] = bool((y == tot * world).all().item())      # two replays: sum, then sum of sums
res["graph_nodes"] = gph.node_count()
gph.destroy()
flags = torch.tensor([int(all(v for k, v in res.items() if k.endswith("_ok") or k.endswith("_exact") or k.endswith("untouched")))])
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
res["ok_all_ranks"] = bool(flags.item())
K.ddp_finalize(local)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
sys.exit(0 if res["ok_all_ranks"] else 1)
