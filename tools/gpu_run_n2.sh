cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cpp_multirank_gpu.py -x -q 2>&1 | tail -30
