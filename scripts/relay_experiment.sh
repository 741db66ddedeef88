# usage (8 GPUs, under gpurun --gpus 8): bash scripts/relay_experiment.sh
# Same box: the host->device ceiling, then the e2e loop direct, with GPUs 0-3 relaying through 4-7, and the reverse.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 20 --warmup 5 --skip-cpu --skip-configs"
timeout 120 $TR --master-port 29551 scripts/h2d_probe.py > gpurun_out/relx_probe.json 2> gpurun_out/relx_probe.err
timeout 200 $TR --master-port 29552 $B --relay off > gpurun_out/relx_off.json 2> gpurun_out/relx_off.err
GAD_RELAY_PLAN="0:4:0.27,1:5:0.27,2:6:0.27,3:7:0.27" timeout 200 $TR --master-port 29553 $B --relay auto > gpurun_out/relx_lo.json 2> gpurun_out/relx_lo.err
GAD_RELAY_PLAN="4:0:0.27,5:1:0.27,6:2:0.27,7:3:0.27" timeout 200 $TR --master-port 29554 $B --relay auto > gpurun_out/relx_hi.json 2> gpurun_out/relx_hi.err
tail -c 200 gpurun_out/relx_*.err | tail -12
