set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tile_gpu.py tests/test_gridlayout_golden.py tests/test_configs_gpu.py tests/test_solver_gpu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -3 gpurun_out/bench_r2d.err; cat gpurun_out/bench_r2d.json | cut -c1-6000
