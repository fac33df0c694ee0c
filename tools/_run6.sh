set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err; echo rc=$?; tail -5 gpurun_out/bench_r2c_n2.err; cut -c1-3000 gpurun_out/bench_r2c_n2.json
timeout 900 python -m pytest tests/test_peer_halo_gpu.py tests/test_amr_gpu.py tests/test_simulator_gpu.py -q -x 2>&1 | tail -5
for k in 1 2 3 4; do timeout 600 python bench.py --config $k --steps 10 --warmup 3 > gpurun_out/bench_r2_c$k.json 2> gpurun_out/bench_r2_c$k.err; echo rc=$?; tail -2 gpurun_out/bench_r2_c$k.err; cut -c1-400 gpurun_out/bench_r2_c$k.json; done
