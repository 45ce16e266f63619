# last evidence run of round 2 on HEAD (1 GPU, ordered by priority: the command is clamped to the GPU minutes that are left)
V=v37
mkdir -p gpurun_out
timeout 240 python tools/parity_report.py --out gpurun_out/r02_parity_report.txt > gpurun_out/r02_parity_report_$V.log 2>&1; echo "parity_report rc=$?"
tail -4 gpurun_out/r02_parity_report.txt
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_all_$V.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t_all_$V.log
tail -3 gpurun_out/r02_t_all_$V.log
timeout 240 python bench.py > gpurun_out/r02_bench_b16_$V.json 2> gpurun_out/r02_bench_b16_$V.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_b16_$V.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
except Exception as e: print("ERR", e)
PY
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:k_conv_haloILi.ELb1E -c 2 -o /tmp/halo_spade -f python tools/profile_step.py --batch 16 > gpurun_out/r02_ncu_$V.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/halo_spade.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_halo_spade_${V}_raw.csv 2>/dev/null
ncu -i /tmp/halo_spade.ncu-rep --page details > gpurun_out/r02_ncu_full_halo_spade_${V}_details.txt 2>/dev/null
ls -la gpurun_out | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
