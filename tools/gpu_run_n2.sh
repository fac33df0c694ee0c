cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i -E "numa|socket|^CPU\(s\)"; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_r2l_n2.json 2> gpurun_out/bench_r2l_n2.err; echo rc=$?
tail -3 gpurun_out/bench_r2l_n2.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2l_n2.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['parity_check']); print(d['roofline_other'].get('kernel_ms_per_step')); print(d['e2e']['value']/1e9, d['e2e']['staging_ms_per_rank'], d['e2e']['host_placement'])"
