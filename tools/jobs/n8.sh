mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_n1_box8_v36.json 2> gpurun_out/r02_bench_n8_v36.err
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --no-cpu-baseline > gpurun_out/r02_bench_n${n}_v36.json 2>> gpurun_out/r02_bench_n8_v36.err
done
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n1_box8_v36.json","gpurun_out/r02_bench_n4_v36.json","gpurun_out/r02_bench_n8_v36.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"])
    except Exception as e: print(f, "ERR", e)
PY
