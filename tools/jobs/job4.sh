set -x
mkdir -p gpurun_out
timeout 600 python tools/ablate_halo.py 2>&1 | tee gpurun_out/r02_ablate_halo_v13.txt
