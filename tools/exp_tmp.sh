timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r01_v5.log 2>&1 || exit 1
grep -o '"value": [0-9.]*' gpurun_out/bench_r01_v5.log | head -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/ncu_launches_v5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused_smooth_kernel|trimv_kernel" -s 48 -c 16 -o gpurun_out/prof_r01_v5_cycle -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/ncu_full_v5.log 2>&1
ls -la gpurun_out/prof_r01_v5_cycle.ncu-rep
