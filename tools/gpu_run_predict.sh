cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predict_gpu.py -q -x 2>&1 | tail -8
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_predict_gpu.py -q -x -k "same_fields or failed_plans or move_two" > gpurun_out/san_predict.log 2>&1; echo san_rc=$?
tail -5 gpurun_out/san_predict.log; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/san_predict.log
