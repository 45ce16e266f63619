# final run of the round on HEAD (v40: k_wgrad_tma split selection): every GPU test, then the bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_v40.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v40.txt
tail -2 gpurun_out/r02_gpu_tests_v40.txt
timeout 150 python bench.py > gpurun_out/r02_bench_b16_v40.json 2> gpurun_out/r02_bench_b16_v40.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_b16_v40.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])"
