mkdir -p gpurun_out
for z in 0 1 2; do echo "== RD_B200_HALO_ZPF=$z"; RD_B200_HALO_ZPF=$z timeout 300 python tools/bench_conv.py --only "gamma" 2>&1 | grep SPADE; done
