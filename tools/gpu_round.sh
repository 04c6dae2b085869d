set -x
V=${V:-v4}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_$V.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r01_$V.log
timeout 600 python bench.py > gpurun_out/bench_r01_n1_$V.json 2> gpurun_out/bench_r01_n1_$V.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r01_n1_$V.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r01_n1_$V.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline","parity")}); print(d["stage_ms"]); print(d["cpu_baseline"])
PY
if [ -n "$BIG" ]; then
timeout 900 python bench.py --config cfg3 --steps 5 --warmup 3 > gpurun_out/bench_r01_cfg3_$V.json 2> gpurun_out/bench_r01_cfg3_$V.err; echo "bench cfg3 rc=$?"
tail -3 gpurun_out/bench_r01_cfg3_$V.err; head -c 3000 gpurun_out/bench_r01_cfg3_$V.json
timeout 600 python bench.py --kmer-size 63 --steps 10 --warmup 3 > gpurun_out/bench_r01_k63_$V.json 2> gpurun_out/bench_r01_k63_$V.err; echo "bench k63 rc=$?"
tail -3 gpurun_out/bench_r01_k63_$V.err; head -c 3000 gpurun_out/bench_r01_k63_$V.json
fi
if [ -n "$NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_$V.csv python tools/prof_one.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'count_kernel|features_kernel|superkmer_kernel|critical_kernel|scatter_kernel' -c 12 -o gpurun_out/prof_r01_$V -f python tools/prof_one.py 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
fi
