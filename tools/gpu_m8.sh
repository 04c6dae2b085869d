# 8-GPU round: cfg2 with --verify (outputs identical to one GPU on the union of the inputs), cfg3 timed without and with phase syncs
V=${V:-r02_m8} N=8 NOTEST=1 CFG="cfg2" BENCH_ARGS="--verify --no-cpu" PHASES= bash tools/gpu_multi.sh
V=${V:-r02_m8} N=8 NOTEST=1 CFG="cfg3" BENCH_ARGS="--no-cpu" PHASES= STEPS=5 bash tools/gpu_multi.sh
cp gpurun_out/bench_${V:-r02_m8}_cfg3_n8.json gpurun_out/bench_${V:-r02_m8}_cfg3_n8_nophase.json
V=${V:-r02_m8} N=8 NOTEST=1 CFG="cfg3" BENCH_ARGS="--no-cpu" PHASES=1 STEPS=3 bash tools/gpu_multi.sh
