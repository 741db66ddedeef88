# usage: bash scripts/sanitize_r2.sh   (GPU box) -- compute-sanitizer over the kernels written / rewritten in round 2
set -x
SAN="compute-sanitizer --error-exitcode 9 --print-limit 5"
$SAN --tool memcheck python -m pytest tests/test_fem2d_gpu.py -m gpu -q -x -k "jitter_n6 or pde_loss" 2>&1 | tail -6
$SAN --tool racecheck --racecheck-report hazard python -m pytest tests/test_fem2d_gpu.py -m gpu -q -x -k "jitter_n5" 2>&1 | tail -8
$SAN --tool memcheck python -m pytest tests/test_glob_cnn.py tests/test_glob_features.py -m gpu -q -x -k "6x6 or 21_b4 or 1d_15x3 or md0" 2>&1 | tail -6
$SAN --tool racecheck --racecheck-report hazard python -m pytest tests/test_glob_cnn.py -m gpu -q -x -k "9x9 and False" 2>&1 | tail -8
$SAN --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rk4_backward and (mesh_dims2 or mesh_dims3)" 2>&1 | tail -6
$SAN --tool racecheck --racecheck-report hazard python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rk4_backward and mesh_dims2" 2>&1 | tail -8
$SAN --tool memcheck python -m pytest tests/test_gpu_peer_loopback.py tests/test_gpu_ell.py -m gpu -q -x -k "world4 or 4-2 or shared_topology_graph_equals and mesh_dims1 or train_batch" 2>&1 | tail -6
$SAN --tool racecheck --racecheck-report hazard python -m pytest tests/test_gpu_peer_loopback.py -m gpu -q -x -k "8-5" 2>&1 | tail -8
