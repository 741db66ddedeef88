python -m pytest tests/test_gpu_ell.py -q > gpurun_out/t4.log 2>&1; tail -5 gpurun_out/t4.log
python -m pytest tests/test_gpu_parity.py -q > gpurun_out/t5.log 2>&1; tail -5 gpurun_out/t5.log
python scripts/kbench.py --check > gpurun_out/kb4.log 2>&1; tail -1 gpurun_out/kb4.log
for t in 192 256 320 480 640; do GAD_ELL_THREADS=$t python scripts/kbench.py --tag thr$t 2>&1 | tail -1; done
GAD_ELL_SMEM=0 python scripts/kbench.py --tag ells0 2>&1 | tail -1
