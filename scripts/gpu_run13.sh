P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ("tag","fwd_us","bwd_us","train_us","step_us","epoch_step_us")})'
python scripts/kbench.py --ring 8 --iters 200 --tag default 2>&1 | tail -1 | python -c "$P"
GAD_LIB=$PWD/g_adaptivity_b200/libv320.so GAD_ELL_THREADS=320 python scripts/kbench.py --ring 8 --iters 200 --tag v320 2>&1 | tail -1 | python -c "$P"
GAD_ELL_THREADS=512 python scripts/kbench.py --ring 8 --iters 200 --tag t512 2>&1 | tail -1 | python -c "$P"
GAD_ELL_THREADS=480 python scripts/kbench.py --ring 8 --iters 200 --tag t480 2>&1 | tail -1 | python -c "$P"
