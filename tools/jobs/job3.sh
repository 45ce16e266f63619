set -x
mkdir -p gpurun_out
export RD_B200_SPADE_BWD_FUSED=0
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k_conv_halo' -c 12 -o /tmp/halo_fwd -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_v12a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k_wgrad_halo' -c 6 -o /tmp/halo_wgrad -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_v12b.log 2>&1
ls -la /tmp/*.ncu-rep
for f in halo_fwd halo_wgrad; do
  ncu -i /tmp/$f.ncu-rep --page raw --csv > gpurun_out/r02_full_${f}_v12_raw.csv 2>/dev/null
  ncu -i /tmp/$f.ncu-rep --page details > gpurun_out/r02_full_${f}_v12_details.txt 2>/dev/null
done
sz=$(du -cm /tmp/halo_fwd.ncu-rep /tmp/halo_wgrad.ncu-rep | tail -1 | cut -f1)
if [ "$sz" -lt 45 ]; then cp /tmp/halo_fwd.ncu-rep /tmp/halo_wgrad.ncu-rep gpurun_out/; else cp /tmp/halo_wgrad.ncu-rep gpurun_out/; fi
du -sh gpurun_out
