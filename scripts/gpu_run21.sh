for c in 256 1024 2048 4096; do echo "chunk $c"; GAD_WIDE_CHUNK=$c python scripts/widebench.py --steps 40 --graph 2>&1 | tail -1; done
GAD_WIDE_CHUNK=2048 python scripts/widebench.py --mesh 200 200 --batch 16 --steps 40 --graph 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
