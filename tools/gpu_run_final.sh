cd /root/repo
timeout 900 python -m pytest tests/test_predict_gpu.py tests/test_cpp_solver.py -q -x 2>&1 | tail -2
