python -m pytest tests -m gpu -q > gpurun_out/t6.log 2>&1; tail -4 gpurun_out/t6.log
python scripts/kbench.py --check 2>&1 | tail -1
python scripts/kbench.py --no-step --iters 30 > gpurun_out/kb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ell_train -s 6 -c 2 -o gpurun_out/prof_r1_v4_train python scripts/kbench.py --no-step --iters 30 > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
