python -m pytest tests/test_gpu_ell.py -q -x > gpurun_out/t7.log 2>&1; tail -12 gpurun_out/t7.log
python scripts/kbench.py --check 2>&1 | tail -1
