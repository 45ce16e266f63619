mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_step_gpu.py -m gpu -x -q -k "full_reference" 2>&1 | tail -4
timeout 200 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_b16_v38.json 2> gpurun_out/r02_bench_b16_v38.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_b16_v38.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['traffic_source'][:40])"
