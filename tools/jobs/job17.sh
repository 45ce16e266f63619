mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r02_bench_b16_v27.json 2> gpurun_out/r02_bench_b16_v27.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_b16_v27.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['losses']['all'])"
