timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/bench_conv.py --only "sp" 2>&1 | grep -E "^sp[56]" | grep -v SPADE
