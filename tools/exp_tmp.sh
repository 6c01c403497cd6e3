timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q11.log 2>&1; echo "rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_q11.log | head -1) illegal=$(grep -c illegal gpurun_out/bench_q11.log)"
