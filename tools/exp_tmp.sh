for t in 40,46,24,8,36 36,46,24,8,40 40,46,24,8,32 32,46,24,8,32 40,36,24,8,40 40,46,16,8,40; do
  MGMC_TILE_ROWS=$t python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-batched > gpurun_out/bench_t3_$t.log 2>&1
  echo "$t $(grep -o '"value": [0-9.]*' gpurun_out/bench_t3_$t.log | head -1)"
done
