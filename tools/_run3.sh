set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_tile_gpu.py -x -q -k "plan_then_scatter and 3-2" 2>&1 | grep -v "^$" | head -60 > gpurun_out/run3_san.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect "tests/test_tile_gpu.py::test_push_deposit_plan_then_scatter[3-2]" --deselect "tests/test_tile_gpu.py::test_push_deposit_plan_then_scatter[3-3]" 2>&1 | tail -15 > gpurun_out/run3_gpu.log
cat gpurun_out/run3_san.log | tail -40
cat gpurun_out/run3_gpu.log
