timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/run_c4_strips.py > gpurun_out/c4_n1b.log 2>&1; tail -1 gpurun_out/c4_n1b.log
