# One GPU-box round: parity tests, the bench line (both arms), optional ncu captures. Outputs under gpurun_out/.
set -x
V=${V:-r02_v1}
CFG=${CFG:-cfg3}
if [ -z "$NOTEST" ]; then
timeout ${TEST_TIMEOUT:-600} python -m pytest tests -m gpu -q --timeout=${PER_TEST_TIMEOUT:-200} --timeout-method=thread ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_$V.log 2>&1; echo "pytest rc=$?"
RC=$?
tail -25 gpurun_out/pytest_gpu_$V.log
if grep -qE "failed|Timeout|error" gpurun_out/pytest_gpu_$V.log && [ -z "$BENCH_ANYWAY" ]; then echo "tests not green: bench skipped"; NOBENCH=1; fi
fi
if [ -z "$NOBENCH" ]; then
timeout ${BENCH_TIMEOUT:-500} python bench.py --config $CFG ${BENCH_ARGS:-} > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_$V.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$V.json"))
print({k:d.get(k) for k in ("value","ms_per_step","e2e","parity","parity_full","e2e_from_files")})
print(d["roofline"]); print(d["stage_ms"]); print(d["cpu_baseline"])
for k in d["kernels"]: print("%-60s %8.3f ms  %8.1f GB/s  frac %.3f (%s)" % (k["kernel"][:60], k["ms"], k["achieved_gbs"] or 0, k["frac"] or 0, k.get("bound")))
PY
fi
if [ -n "$REFARM" ]; then
timeout 900 python bench.py --config $CFG --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$V.json 2> gpurun_out/bench_ref_$V.err; echo "ref arm rc=$?"
cat gpurun_out/bench_ref_$V.json | head -c 1500
fi
if [ -n "$NCU" ]; then
# launch list (gpu__time_duration only) of two finds, then ONE full capture of every kernel of one find; the .ncu-rep stays on the
# box (it exceeds what gpurun copies back): its summaries are written here as text
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${V}_$CFG.csv python tools/prof_one.py 2 1.0 ${KM:-31} $CFG > gpurun_out/ncu_launches_$V.log 2>&1; echo "ncu launches rc=$?"
tail -2 gpurun_out/ncu_launches_$V.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-count_kernel|features_kernel|superkmer_kernel|critical_kernel|scatter_kernel|table_build_kernel|bloom_neighbor_insert_kernel|mphf_level_kernel}" -c ${NCU_COUNT:-24} -o /tmp/prof_${V}_$CFG -f python tools/prof_one.py 1 1.0 ${KM:-31} $CFG > gpurun_out/ncu_full_$V.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_$V.log
python tools/ncu_summary.py /tmp/prof_${V}_$CFG.ncu-rep > gpurun_out/ncu_full_${V}_$CFG.md 2>&1
cp profiles/ncu_traffic.json /tmp/ncu_traffic_before.json 2>/dev/null
python tools/ncu_traffic.py /tmp/prof_${V}_$CFG.ncu-rep ${CFG}_k${KM:-31} > gpurun_out/ncu_traffic_${V}.log 2>&1; cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_${V}.json
for kern in ${NCU_LINES:-count_kernel superkmer_kernel critical_kernel table_build_kernel features_kernel scatter_kernel}; do
python tools/ncu_lines.py /tmp/prof_${V}_$CFG.ncu-rep $kern 0 40 > gpurun_out/ncu_lines_${kern}_${V}.txt 2>&1
done
ls -la /tmp/prof_${V}_$CFG.ncu-rep; head -30 gpurun_out/ncu_full_${V}_$CFG.md | cut -c1-250
fi
