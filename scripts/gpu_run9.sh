python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; tail -3 gpurun_out/t10.log
python scripts/ktrace.py 2>&1 | tail -13
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ("tag","fwd_us","bwd_us","train_us","step_us","epoch_step_us")})'
python scripts/kbench.py --ring 8 --iters 200 --tag new 2>&1 | tail -1 | python -c "$P"
