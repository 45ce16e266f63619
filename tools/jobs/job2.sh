set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_tc_gpu.py -m gpu -x -q -k "spade or norm" > gpurun_out/r02_t_spade_v12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t_spade_v12.log
tail -5 gpurun_out/r02_t_spade_v12.log
timeout 300 python tools/bench_spade_bwd.py 2>&1 | tee gpurun_out/r02_bench_spade_bwd_v12.txt
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k_wgrad_halo|k_conv_halo' -c 36 -o gpurun_out/r02_full_halo_v11 -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_v11.log 2>&1
ls -la gpurun_out/*.ncu-rep
