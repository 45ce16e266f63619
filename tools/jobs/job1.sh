set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_all_v11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t_all_v11.log
timeout 300 python bench.py > gpurun_out/r02_bench_b16_v11.json 2> gpurun_out/r02_bench_b16_v11.err
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k_wgrad_halo|k_conv_halo' -c 36 -o gpurun_out/r02_full_halo_v11 -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_v11.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r02_t_all_v11.log; cat gpurun_out/r02_bench_b16_v11.json
