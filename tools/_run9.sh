set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cpp_solver.py tests/test_cpp_mirror.py tests/test_tile_gpu.py tests/test_configs_gpu.py tests/test_solver_gpu.py -x -q 2>&1 | tail -6
