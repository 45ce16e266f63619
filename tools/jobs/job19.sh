mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q 2>&1 | tail -4
for sw in 0 1; do echo "== RD_B200_WGH_DYSW=$sw"; RD_B200_WGH_DYSW=$sw timeout 300 python tools/bench_conv.py --only "sp" 2>&1 | grep -E "^sp[56]" | grep -v SPADE; done
