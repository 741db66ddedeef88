python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; tail -5 gpurun_out/t10.log
python scripts/ktrace.py 2>&1 | tail -14
GAD_TAIL_CTA=0 python scripts/ktrace.py 2>&1 | tail -2
python scripts/kbench.py --ring 8 --iters 200 2>&1 | tail -1
GAD_TAIL_CTA=0 python scripts/kbench.py --ring 8 --iters 200 2>&1 | tail -1
