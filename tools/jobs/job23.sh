for w in 0 1 3 4; do
  RD_B200_WGRAD_WAVES=$w timeout 200 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('waves=$w', d['value'], d['ms_per_step'])"
done
