# k_wgrad_tma split selection (wave quantisation): correctness + A/B on one box
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q 2>&1 | tail -2
RD_B200_WGRAD_TRACE=1 timeout 200 python -m pytest tests/test_step_gpu.py tests/test_fullsize_properties_gpu.py -m gpu -x -q 2> gpurun_out/r02_wgrad_split_trace_tests.txt | tail -2
for b in 0 1; do
  RD_B200_WGRAD_BALANCE=$b RD_B200_WGRAD_TRACE=1 timeout 90 python tools/bench_conv.py > gpurun_out/r02_bench_conv_balance$b.txt 2> gpurun_out/r02_wgrad_split_trace_$b.txt
  RD_B200_WGRAD_BALANCE=$b timeout 120 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('balance=$b', d['value'], d['ms_per_step'], d['e2e']['value'], d['losses'])"
done
paste -d'\n' gpurun_out/r02_bench_conv_balance0.txt gpurun_out/r02_bench_conv_balance1.txt | grep -v SPADE | sed -e 's/.*| wgrad/wgrad/' | paste - - | head -24
