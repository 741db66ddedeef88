for c in 4 8 16; do echo "100x100 C>=$c"; GAD_CLUSTER_MIN=$c python scripts/widebench.py --steps 40 --graph 2>&1 | tail -1; done
for c in 2 4 8; do echo "60x60 C>=$c"; GAD_CLUSTER_MIN=$c python scripts/widebench.py --mesh 60 60 --batch 256 --steps 40 --graph 2>&1 | tail -1; done
