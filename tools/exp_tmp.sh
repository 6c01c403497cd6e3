for v in b200 lu3; do
for t in 40,46,24,8,44 38,46,24,8,42 38,38,24,8,42; do
  lib=build/libmgmc_$v.so; [ $v = b200 ] && lib=multigridmc_b200/csrc/libmgmc_b200.so
  MGMC_LIB=$lib MGMC_TILE_ROWS=$t python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lu_${v}_$t.log 2>&1
  echo "$v $t $(grep -o '"value": [0-9.]*' gpurun_out/bench_lu_${v}_$t.log | head -1) $(grep -o '"qoi_mean": [-0-9.e]*' gpurun_out/bench_lu_${v}_$t.log)"
done
done
