# usage: bash scripts/summarize_prof_r2.sh <tag> <out-prefix>   (build container; reads the reports gpurun brought back)
# One command turns the captures of scripts/gpu_prof_r2.sh into everything under profiles/ that bench.py and the
# docs cite -- the ncu summaries, the launch list and profiles/traffic.json (roofline.traffic) -- so they cannot drift.
TAG=${1:-r02v2}
OUT=${2:-r02_v2}
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_bench.ncu-rep k_ell_train > profiles/${OUT}_ncu_k_ell_train.md
python scripts/ncu_traffic.py gpurun_out/prof_${TAG}_bench.ncu-rep > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_fem2d.ncu-rep k_fem2d > profiles/${OUT}_ncu_fem2d.md
python scripts/summarize_launches.py gpurun_out/launches_${TAG}.csv "ncu launch list of: python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-configs" > profiles/${OUT}_launches.md
cp gpurun_out/launches_${TAG}.csv profiles/${OUT}_launches.csv
