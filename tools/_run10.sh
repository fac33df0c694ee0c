set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cpp_solver.py tests/test_host_api_gpu.py tests/test_peer_halo_gpu.py -x -q 2>&1 | tail -8
for k in 1 2 3 4 5; do timeout 600 python bench.py --config $k --host cpp --steps 10 --warmup 3 > gpurun_out/bench_r2_cpp_c$k.json 2> gpurun_out/bench_r2_cpp_c$k.err; echo rc=$?; tail -2 gpurun_out/bench_r2_cpp_c$k.err; cut -c1-330 gpurun_out/bench_r2_cpp_c$k.json; done
