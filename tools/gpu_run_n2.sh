cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cpp_multirank_gpu.py tests/test_peer_halo_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2z_n2.json 2> gpurun_out/bench_r2z_n2.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2z_n2.json') if l.startswith('{')][0]
print(d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['python_host']['ms_per_step'],3), round(d['e2e']['value']/1e9,2), d['parity_check']['ok'], d['parity_check_cpp_host']['ok'], d['clocks'])"
