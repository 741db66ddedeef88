# usage: bash scripts/gpu_bench.sh <tag>   -- bench, then the ncu launch list and one full capture of the same command
TAG=${1:-r01}
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --ring 2"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ell_train -s 30 -c 2 -o gpurun_out/prof_${TAG}_bench $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
tail -2 gpurun_out/ncu_f_$TAG.log
