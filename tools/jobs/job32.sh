# sanity of the rebuilt library (wg_geometry refactor + rd_wgrad_tma_plan): convolution tests, short bench
timeout 30 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q 2>&1 | tail -1
timeout 30 python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['losses']['all'])"
