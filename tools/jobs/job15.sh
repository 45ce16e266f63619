mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r02_bench_b16_v24.json 2> gpurun_out/r02_bench_b16_v24.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_b16_v24.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['losses']['all'], d['roofline']['frac'])"
