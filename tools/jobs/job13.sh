mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:k_conv_haloILi.ELb1E -c 2 -o /tmp/halo_spade -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_v22.log 2>&1
ls -la /tmp/halo_spade.ncu-rep
cp /tmp/halo_spade.ncu-rep gpurun_out/r02_full_halo_spade_v22.ncu-rep
