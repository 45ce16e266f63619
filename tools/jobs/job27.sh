# launch list of one eager training iteration with the tensor pipe's active share per launch ("ncu tensor-pipe % quoted per kernel")
V=v37
mkdir -p gpurun_out
RD_B200_TRACE_CONV=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r02_launches_step_b16_$V.csv python tools/profile_step.py --batch 16 > gpurun_out/r02_trace_conv_$V.txt 2>&1; echo "ncu rc=$?"
python tools/conv_table.py gpurun_out/r02_launches_step_b16_$V.csv gpurun_out/r02_trace_conv_$V.txt > gpurun_out/r02_conv_table_step_b16_$V.txt 2>&1; echo "table rc=$?"
python tools/summarize_launches_bw.py gpurun_out/r02_launches_step_b16_$V.csv 60 > gpurun_out/r02_launches_step_b16_$V.txt 2>&1
head -12 gpurun_out/r02_conv_table_step_b16_$V.txt
tail -2 gpurun_out/r02_trace_conv_$V.txt | cut -c1-300
