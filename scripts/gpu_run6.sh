python -m pytest tests/test_gpu_ell.py tests/test_gpu_parity.py -q -x > gpurun_out/t8.log 2>&1; tail -6 gpurun_out/t8.log
python scripts/kbench.py --check 2>&1 | tail -1
