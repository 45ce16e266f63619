mkdir -p gpurun_out
timeout 200 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_b16_v39.json 2> gpurun_out/r02_bench_b16_v39.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_b16_v39.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['dominant_kernel']['fused_in_step'])"
# ncu --set full of the largest single launch of the step: k_wgrad_tma on the sp4 gamma|beta weight gradient (7th k_wgrad_tma launch of the iteration)
timeout 240 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_wgrad_tma -c 7 -o /tmp/wgrad_tma -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_v39.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/wgrad_tma.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_wgrad_tma_v39_raw.csv 2>/dev/null
ncu -i /tmp/wgrad_tma.ncu-rep --page details > gpurun_out/r02_ncu_full_wgrad_tma_v39_details.txt 2>/dev/null
ls -la gpurun_out | tail -5
