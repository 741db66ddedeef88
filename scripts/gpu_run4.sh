for t in 160 192 224 256; do GAD_LIB=$PWD/g_adaptivity_b200/libgad_m512.so GAD_ELL_THREADS=$t python scripts/kbench.py --tag m512_t$t 2>&1 | tail -1; done
for t in 224 320; do GAD_ELL_THREADS=$t python scripts/kbench.py --tag m640_t$t 2>&1 | tail -1; done
