python -m pytest tests/test_gpu_invariance.py -x -q 2>&1 | tail -15
python profiles/run_configs.py > gpurun_out/configs_r01_v4.log 2>&1; cat gpurun_out/configs_r01_v4.log | tail -8
