python -m pytest tests/test_gpu_parity.py tests/test_gpu_ell.py -x -q > gpurun_out/t10.log 2>&1; tail -2 gpurun_out/t10.log
python scripts/widebench.py --steps 40 --graph 2>&1 | tail -1
python scripts/widebench.py --mesh 200 200 --batch 16 --steps 40 --graph 2>&1 | tail -1
python scripts/widebench.py --mesh 60 60 --batch 256 --steps 40 --graph 2>&1 | tail -1
