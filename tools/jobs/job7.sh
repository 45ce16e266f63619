mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_all_v16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t_all_v16.log
tail -3 gpurun_out/r02_t_all_v16.log
RD_B200_HALO_DUAL=0 timeout 300 python bench.py > gpurun_out/r02_bench_b16_v16_dual0.json 2> gpurun_out/r02_bench_b16_v16_dual0.err
timeout 300 python bench.py > gpurun_out/r02_bench_b16_v16.json 2> gpurun_out/r02_bench_b16_v16.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_b16_v16_dual0.json","gpurun_out/r02_bench_b16_v16.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["losses"]["all"])
    except Exception as e: print(f, "ERR", e)
PY
