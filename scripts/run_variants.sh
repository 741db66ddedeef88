#!/bin/bash
# usage: scripts/run_variants.sh "<kbench args>" lib1 lib2 ...   (run on the GPU box)
ARGS="$1"; shift
for lib in "$@"; do
  GAD_LIB=$PWD/g_adaptivity_b200/$lib python scripts/kbench.py $ARGS --tag $lib 2>&1 | tail -1
done
