mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_parity.py > gpurun_out/r02_ddp_parity_n2_v30.log 2>&1
tail -3 gpurun_out/r02_ddp_parity_n2_v30.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox_v30.json 2> gpurun_out/r02_bench_n2_v30.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02_bench_n2_v30.json 2>> gpurun_out/r02_bench_n2_v30.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n1_samebox_v30.json","gpurun_out/r02_bench_n2_v30.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"])
    except Exception as e: print(f, "ERR", e)
PY
