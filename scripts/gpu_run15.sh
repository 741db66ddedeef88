python -m pytest tests/test_gpu_ell.py tests/test_gpu_parity.py -x -q > gpurun_out/t10.log 2>&1; tail -2 gpurun_out/t10.log
python bench.py --steps 1000 --warmup 20 --skip-cpu > gpurun_out/b15.json 2> gpurun_out/b15.err; grep -a "^{" gpurun_out/b15.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -2 gpurun_out/b15.err
