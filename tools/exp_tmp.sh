for k in 1 2 3; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_rep_$k.log 2>&1; echo "rep $k rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_rep_$k.log | head -1) illegal=$(grep -c 'illegal' gpurun_out/bench_rep_$k.log)"
done
python bench.py > gpurun_out/bench_rep_default.log 2>&1; echo "default rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_rep_default.log | head -1) illegal=$(grep -c 'illegal' gpurun_out/bench_rep_default.log)"
