N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 280 python scripts/cfg5_sweep.py > gpurun_out/cfg5_$N.json 2> gpurun_out/cfg5_$N.err
else
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/cfg5_sweep.py > gpurun_out/cfg5_$N.json 2> gpurun_out/cfg5_$N.err
fi
echo "rc=$?"; grep -a "^{" gpurun_out/cfg5_$N.json; tail -3 gpurun_out/cfg5_$N.err
