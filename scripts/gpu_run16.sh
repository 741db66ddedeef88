P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ("tag","fwd_us","bwd_us","train_us","step_us","epoch_step_us","fwd_err","grad_err")})'
python scripts/kbench.py --ring 8 --iters 200 --tag default --check 2>&1 | tail -1 | python -c "$P"
GAD_LIB=$PWD/g_adaptivity_b200/libffma2.so python scripts/kbench.py --ring 8 --iters 200 --tag ffma2 --check 2>&1 | tail -1 | python -c "$P"
