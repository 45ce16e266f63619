mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "spade" 2>&1 | tail -3
timeout 300 python tools/bench_spade_bwd.py 2>&1
