mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_all_v19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t_all_v19.log
tail -3 gpurun_out/r02_t_all_v19.log
RD_B200_HALO_DUAL=0 RD_B200_TMA_DUAL=0 timeout 300 python bench.py > gpurun_out/r02_bench_b16_v19_dual0.json 2> gpurun_out/r02_bench_b16_v19_dual0.err
timeout 300 python bench.py > gpurun_out/r02_bench_b16_v19.json 2> gpurun_out/r02_bench_b16_v19.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_b16_v19_dual0.json","gpurun_out/r02_bench_b16_v19.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["losses"]["all"])
    except Exception as e: print(f, "ERR", e)
PY
RD_B200_TRACE_CONV=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b16_v19.csv python tools/profile_step.py --batch 16 > gpurun_out/r02_trace_conv_v19.txt 2>&1
tail -2 gpurun_out/r02_trace_conv_v19.txt
