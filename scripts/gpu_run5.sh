python scripts/kbench.py --no-step --iters 30 > gpurun_out/kb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ell_train" -s 12 -c 2 -o gpurun_out/prof_r1_v5_train python scripts/kbench.py --no-step --iters 30 > gpurun_out/ncu5.log 2>&1
tail -3 gpurun_out/ncu5.log
