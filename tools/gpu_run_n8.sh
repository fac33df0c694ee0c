cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
lscpu | grep -i -E "numa|socket|^CPU\(s\)"; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2o_n8.json 2> gpurun_out/bench_r2o_n8.err; echo rc=$?
tail -3 gpurun_out/bench_r2o_n8.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2o_n8.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['parity_check']['ok'], d['parity_check']['max_rel'], d['clocks']); print(d['roofline_other'].get('kernel_ms_per_step')); print(d['e2e']['value']/1e9, d['e2e']['staging_ms_per_rank'], d['e2e']['host_placement'][0])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r2o_c4_n8.json 2> gpurun_out/bench_r2o_c4_n8.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2o_c4_n8.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['parity_check']['ok'], d['parity_check']['max_rel']); print(d['roofline_other'].get('kernel_ms_per_step')); print(d['e2e']['value']/1e9)"
