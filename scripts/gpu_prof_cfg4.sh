# usage: bash scripts/gpu_prof_cfg4.sh   (on the GPU box, under gpurun)
# cfg 4 (one 200x200 mesh, 64 RK4 steps, forward): plain run, ncu launch list, one `--set full` capture of
# k_wide_persist.  The kernel polls other CTAs: if a profiler pass ever serialised its CTAs the 2 s poll bound traps
# (a CUDA error, not a hang), and every command runs under its own timeout.
CMD="python scripts/cfg4_bench.py 200"
timeout 200 $CMD > gpurun_out/plain_cfg4.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_cfg4.log; exit 1; }
tail -2 gpurun_out/plain_cfg4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg4.csv $CMD > gpurun_out/ncu_l_cfg4.log 2>&1
grep -c k_wide_persist gpurun_out/launches_cfg4.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wide_persist -s 3 -c 1 -o gpurun_out/prof_cfg4_persist $CMD > gpurun_out/ncu_f_cfg4.log 2>&1
tail -3 gpurun_out/ncu_f_cfg4.log
ls -la gpurun_out/*cfg4*
