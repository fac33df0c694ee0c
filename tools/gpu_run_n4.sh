cd /root/repo
mkdir -p gpurun_out
for host in python cpp; do
extra=""; [ $host = cpp ] && extra="--no-e2e"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --host $host $extra > gpurun_out/bench_r2r_${host}_n4.json 2> gpurun_out/bench_r2r_${host}_n4.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2r_${host}_n4.json') if l.startswith('{')][0]
print('$host', d['value']/1e9, d['ms_per_step'], d['parity_check']['ok'], d['parity_check']['max_rel'], d['clocks'], (d.get('e2e') or {}).get('value'))"
done
