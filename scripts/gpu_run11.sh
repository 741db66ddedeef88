timeout 280 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/t11.log 2>&1; tail -5 gpurun_out/t11.log
