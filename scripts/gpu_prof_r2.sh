# usage: bash scripts/gpu_prof_r2.sh <tag>   (on the GPU box, under gpurun)
# 1. the bench command without a profiler (must exit 0), 2. its ncu launch list, 3. one `--set full` capture of the
# dominant kernel (k_ell_train) of the same command, 4. one `--set full` capture of the 2-D FEM kernels.
TAG=${1:-r02}
CMD="python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-configs"
timeout 300 $CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ell_train -s 40 -c 2 -o gpurun_out/prof_${TAG}_bench $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
tail -2 gpurun_out/ncu_f_$TAG.log
timeout 300 python scripts/fem2d_check.py > gpurun_out/plain_fem2d_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fem2d -s 8 -c 2 -o gpurun_out/prof_${TAG}_fem2d python scripts/fem2d_check.py > gpurun_out/ncu_fem_$TAG.log 2>&1
tail -2 gpurun_out/ncu_fem_$TAG.log
ls -la gpurun_out/*.ncu-rep
