cd /root/repo
mkdir -p gpurun_out
for host in cpp python; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e --host $host > gpurun_out/bench_r2p_${host}_n2.json 2> gpurun_out/bench_r2p_${host}_n2.err; echo rc=$?
tail -2 gpurun_out/bench_r2p_${host}_n2.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2p_${host}_n2.json') if l.startswith('{')][0]
print('$host', d['value']/1e9, d['ms_per_step'], d.get('parity_check'), d['gpu_launches'], d['clocks'])"
done
