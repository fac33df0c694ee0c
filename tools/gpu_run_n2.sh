cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/gputests_2gpu.log; cat gpurun_out/gputests_2gpu.log | tail -6
