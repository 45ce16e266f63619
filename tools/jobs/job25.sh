timeout 600 python -m pytest tests/test_step_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['dominant_kernel']['ms'])"
