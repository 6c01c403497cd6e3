python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q7.log 2>&1; grep -o '"value": [0-9.]*' gpurun_out/bench_q7.log | head -1; grep -o '"qoi_mean": [-0-9.e]*' gpurun_out/bench_q7.log
