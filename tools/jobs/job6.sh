mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_t_conv_v16.log
for du in 0 1; do
  echo "== RD_B200_HALO_DUAL=$du"
  RD_B200_HALO_DUAL=$du timeout 300 python tools/bench_conv.py --only sp 2>&1 | grep -E "^sp[56]"
done 2>&1 | tee gpurun_out/r02_bench_conv_dual_v16.txt
