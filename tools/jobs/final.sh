# round-end evidence run (1 GPU): tests, smoke, bench lines of every workload, per-layer conv table, launch list
V=${1:-v30}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_all_$V.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t_all_$V.log
tail -3 gpurun_out/r02_t_all_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r02_bench_b16_$V.json 2> gpurun_out/r02_bench_b16_$V.err
for wl in brats-dropout ncanda infer-sweep; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02_bench_${wl}_$V.json 2>> gpurun_out/r02_bench_b16_$V.err
done
timeout 300 python bench.py --workload infer-sweep --sweep-dedup --no-cpu-baseline > gpurun_out/r02_bench_infer-sweep-dedup_$V.json 2>> gpurun_out/r02_bench_b16_$V.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_$V.json 2>> gpurun_out/r02_bench_b16_$V.err
timeout 300 python tools/bench_conv.py > gpurun_out/r02_bench_conv_b16_$V.txt 2>&1
RD_B200_TRACE_CONV=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b16_$V.csv python tools/profile_step.py --batch 16 > gpurun_out/r02_trace_conv_$V.txt 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*_$V.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
