python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; tail -3 gpurun_out/t10.log
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ("tag","nodes","tiles","ell","fwd_us","bwd_us","step_us","epoch_step_us","gnodes_per_s","fwd_err","grad_err")})'
python scripts/kbench.py --mesh 100 100 --batch 64 --ring 2 --iters 40 --tag m100wide --check 2>&1 | tail -1 | python -c "$P"
GAD_NO_WIDE=1 python scripts/kbench.py --mesh 100 100 --batch 64 --ring 2 --iters 40 --tag m100csr 2>&1 | tail -1 | python -c "$P"
python scripts/kbench.py --mesh 200 200 --batch 16 --ring 2 --iters 40 --tag m200wide 2>&1 | tail -1 | python -c "$P"
