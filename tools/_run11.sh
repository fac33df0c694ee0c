set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m | head -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_r2e_n8.json 2> gpurun_out/bench_r2e_n8.err; echo rc=$?; tail -3 gpurun_out/bench_r2e_n8.err; cut -c1-200 gpurun_out/bench_r2e_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r2e_c4_n8.json 2> gpurun_out/bench_r2e_c4_n8.err; echo rc=$?; tail -3 gpurun_out/bench_r2e_c4_n8.err; cut -c1-200 gpurun_out/bench_r2e_c4_n8.json
