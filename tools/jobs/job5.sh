mkdir -p gpurun_out
timeout 120 tools/bin/mma_rate mn 2>&1 | tee gpurun_out/r02_mma_rate_mn_v13.txt
for sl in "0,0,0" "50,20,40" "100,40,80" "400,40,80" "200,0,0" "200,100,200"; do
  echo "== RD_B200_HALO_SLEEP=$sl"
  RD_B200_HALO_SLEEP=$sl timeout 200 python tools/bench_conv.py --only sp6 2>&1 | grep -v "^FLOP"
done 2>&1 | tee gpurun_out/r02_halo_sleep_sweep_v13.txt
