python __graft_entry__.py smoke 2>&1 | tail -2
python profiles/run_configs.py > gpurun_out/configs_r01_v5.log 2>&1; grep '^{' gpurun_out/configs_r01_v5.log | cut -c1-230
