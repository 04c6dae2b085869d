set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01_s4.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r01_s4.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r01_n1_v3.json 2> gpurun_out/bench_r01_n1_v3.err; echo "bench rc=$?"
cat gpurun_out/bench_r01_n1_v3.json | head -c 1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref_v3.json 2>&1; echo "ref rc=$?"
cat gpurun_out/bench_r01_ref_v3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v3.csv python tools/prof_one.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'count_kernel|features_kernel|superkmer_kernel|critical_kernel|scatter_kernel|pack_kernel' -c 16 -o gpurun_out/prof_r01_v3 -f python tools/prof_one.py 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
