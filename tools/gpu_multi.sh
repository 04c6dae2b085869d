# N-GPU round (gpurun --gpus N): NCCL parity tests + bench at N ranks with --verify. Outputs under gpurun_out/.
set -x
V=${V:-r02_m1}
N=${N:-2}
CFG=${CFG:-cfg3}
if [ -z "$NOTEST" ]; then
timeout 400 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread -k "two_gpus or dist_path" > gpurun_out/pytest_gpu_multi_$V.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu_multi_$V.log
fi
for C in $CFG; do
MTG_DIST_PHASES=${PHASES:-} timeout ${BENCH_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $C --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${V}_${C}_n$N.json 2> gpurun_out/bench_${V}_${C}_n$N.err; echo "bench $C rc=$?"
tail -4 gpurun_out/bench_${V}_${C}_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${V}_${C}_n$N.json"))
    print("$C N=$N", {k:d.get(k) for k in ("value","ms_per_step","e2e")})
    print({k:round(v,2) for k,v in d["stage_ms"].items()})
except Exception as e: print("no json", e)
PY
done
