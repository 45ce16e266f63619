mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -x -q 2>&1 | tail -4
for nw in 0 1; do echo "== RD_B200_HALO_NARROW=$nw"; RD_B200_HALO_NARROW=$nw timeout 300 python tools/ablate_halo.py --only ch; done
