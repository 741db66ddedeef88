python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; tail -3 gpurun_out/t10.log
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ("tag","nodes","fwd_us","bwd_us","train_us","step_us","epoch_step_us","gnodes_per_s","fwd_err","grad_err")})'
python scripts/kbench.py --ring 8 --iters 200 --tag cfg2 --check 2>&1 | tail -1 | python -c "$P"
python scripts/kbench.py --mesh 50 50 --batch 256 --ring 2 --iters 40 --tag m50 --check 2>&1 | tail -1 | python -c "$P"
python scripts/widebench.py --steps 40 --graph 2>&1 | tail -1
