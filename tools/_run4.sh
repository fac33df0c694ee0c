set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tile_gpu.py tests/test_configs_gpu.py tests/test_step_oracle.py tests/test_solver_gpu.py -x -q 2>&1 | tail -15
cat gpurun_out/parity_configs.json
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
PHB_MB_ONLY=tile PHB_MB_GS=8 timeout 600 python tools/microbench.py c1 c2 c3 c4 2>&1 | grep -v "^$" | tail -24
