python -m pytest tests/test_gpu_ell.py -q -x > gpurun_out/t9.log 2>&1; tail -6 gpurun_out/t9.log
python scripts/kbench.py 2>&1 | tail -1
