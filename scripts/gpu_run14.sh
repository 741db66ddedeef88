python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; tail -3 gpurun_out/t10.log
python scripts/cfgbench.py 2>&1 | tail -5
