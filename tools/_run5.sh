set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tile_gpu.py tests/test_configs_gpu.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -3 gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b.json
PHB_MB_ONLY=tile PHB_MB_GS=8,16 timeout 600 python tools/microbench.py c5 2>&1 | grep -v "^$" | tail -24
