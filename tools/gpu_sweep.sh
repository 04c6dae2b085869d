# A/B of count-stage knobs on one box: `gpurun -- bash tools/gpu_sweep.sh`; V names the output files, VARS the env settings to try.
V=${V:-r02_sweep}
VARS=${VARS:-"MTG_COUNT_NOPAYLOAD=0 MTG_COUNT_NOPAYLOAD=1"}
mkdir -p gpurun_out
[ -n "$NOTEST" ] || timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "count or solid or poly" --timeout=300 --timeout-method=thread 2>&1 | tail -3
for S in $VARS; do
T=$(echo $S | tr "=," "__")
env $(echo $S | tr "," " ") timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --files-steps 0 > gpurun_out/bench_${V}_$T.json 2> gpurun_out/bench_${V}_$T.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${V}_$T.json")); print("$S", d["ms_per_step"], {k:round(v,2) for k,v in d["stage_ms"].items() if k.startswith("count") or k.startswith("graph")}, d["counts"]["nb_multipass_groups"], d["parity_full"]["breakpoints_equal"])
PY
done
