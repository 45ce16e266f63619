mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_t_conv_v17.log
echo "== all dual off (HALO_DUAL=0 TMA_DUAL=0)"; RD_B200_HALO_DUAL=0 RD_B200_TMA_DUAL=0 timeout 300 python tools/bench_conv.py 2>&1 | grep -vE "SPADE:"
echo "== halo NT=2 dual only (HALO_DUAL1=0 TMA_DUAL=0)"; RD_B200_HALO_DUAL1=0 RD_B200_TMA_DUAL=0 timeout 300 python tools/bench_conv.py 2>&1 | grep -vE "SPADE:"
echo "== all dual on (default)"; timeout 300 python tools/bench_conv.py 2>&1 | grep -vE "SPADE:"
echo "== tma dual from n_tile 64"; RD_B200_TMA_DUAL=64 timeout 300 python tools/bench_conv.py 2>&1 | grep -vE "SPADE:"
