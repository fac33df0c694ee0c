# what the driver runs at round end, on one B200 box: the GPU parity suite, the smoke test, both bench arms
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
