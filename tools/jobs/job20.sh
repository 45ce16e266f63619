mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_b16_v32.json 2> gpurun_out/r02_bench_b16_v32.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_b16_v32.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['losses']['all'])"
