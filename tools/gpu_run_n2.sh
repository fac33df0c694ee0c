cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cpp_multirank_gpu.py tests/test_peer_halo_gpu.py tests/test_amr_gpu.py -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_final_n2.json') if l.startswith('{')][0]
print(d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],3), d['parity_check']['ok'], d['parity_check_cpp_host']['ok'])"
