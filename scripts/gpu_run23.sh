for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_ell.py -x -q -k cluster 2>&1 | tail -1; done
python scripts/widebench.py --steps 40 --graph 2>&1 | tail -1
python scripts/widebench.py --mesh 60 60 --batch 256 --steps 40 --graph 2>&1 | tail -1
GAD_CLUSTER_MAX=16 python scripts/widebench.py --mesh 200 200 --batch 16 --steps 40 --graph 2>&1 | tail -1
GAD_CLUSTER_MIN=8 GAD_CLUSTER_MAX=8 python scripts/widebench.py --steps 40 --graph 2>&1 | tail -1
