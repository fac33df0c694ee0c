cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_predict_gpu.py::test_same_fields_every_plan_holds[1-3-eps_0]" -q 2>&1 | grep -E "^E|Error|assert" | head -20
timeout 600 python tools/diff_predict.py 2>&1 | grep -v Warn | tail -30
