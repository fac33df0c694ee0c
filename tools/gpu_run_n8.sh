cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-e2e --host cpp > gpurun_out/bench_r2r_cpp_n8.json 2> gpurun_out/bench_r2r_cpp_n8.err; echo rc=$?
tail -2 gpurun_out/bench_r2r_cpp_n8.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2r_cpp_n8.json') if l.startswith('{')][0]
print('cpp', d['value']/1e9, d['ms_per_step'], d['parity_check']['ok'], d['parity_check']['max_rel'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29584 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 --no-cpu --no-e2e --host cpp > gpurun_out/bench_r2r_cpp_c4_n8.json 2> gpurun_out/bench_r2r_cpp_c4_n8.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2r_cpp_c4_n8.json') if l.startswith('{')][0]
print('cpp c4', d['value']/1e9, d['ms_per_step'], d['parity_check']['ok'], d['parity_check']['max_rel'])"
