V=r02_m8 N=8 NOTEST=1 CFG="cfg2" BENCH_ARGS="--verify --no-cpu" PHASES= bash tools/gpu_multi.sh
V=r02_m8 N=8 NOTEST=1 CFG="cfg3" BENCH_ARGS="--no-cpu" PHASES=1 STEPS=4 bash tools/gpu_multi.sh
