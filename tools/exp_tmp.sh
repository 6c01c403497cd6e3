timeout 300 python -m pytest tests/test_strips.py -x -q -k biharmonic 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/run_c4_strips.py > gpurun_out/c4_n2.log 2>&1; tail -1 gpurun_out/c4_n2.log
