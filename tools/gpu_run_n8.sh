cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_final_n8.json 2> gpurun_out/bench_final_n8.err; echo rc=$?
tail -2 gpurun_out/bench_final_n8.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_final_n8.json') if l.startswith('{')][0]
print(d['host'], d['value']/1e9, d['ms_per_step'], d['python_host']['ms_per_step'], d['parity_check']['ok'], d['parity_check_cpp_host']['ok'], d['clocks'], d.get('cpp_host_error')); print(d['e2e']['value']/1e9, d['e2e']['staging_ms_per_rank'][0], d['e2e']['staging_ms_per_rank'][7])"
