mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_loss_cases_gpu.py tests/test_step_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_recon|k_split|k_lrelu|k_add_n' --csv --log-file gpurun_out/r02_launches_recon_v33.csv python tools/profile_step.py --batch 16 > /dev/null 2>&1
python tools/summarize_launches_bw.py gpurun_out/r02_launches_recon_v33.csv
