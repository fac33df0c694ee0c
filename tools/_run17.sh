set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_peer_halo_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2j_n2.json 2> gpurun_out/bench_r2j_n2.err; echo rc=$?; tail -3 gpurun_out/bench_r2j_n2.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2j_n2.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['parity_check']); print(d['roofline_other'].get('kernel_ms_per_step'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --config 3 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2j_c3_n2.json 2> gpurun_out/bench_r2j_c3_n2.err; echo rc=$?; tail -3 gpurun_out/bench_r2j_c3_n2.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2j_c3_n2.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['parity_check']['ok'], d['parity_check']['max_rel']); print(d['roofline_other'].get('kernel_ms_per_step'))"
