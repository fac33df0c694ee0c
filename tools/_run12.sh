set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tile_gpu.py tests/test_gpu_parity.py tests/test_configs_gpu.py tests/test_solver_gpu.py tests/test_cpp_mirror.py -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -3 gpurun_out/bench_r2f.err; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2f.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9); print(d['roofline']); print(d['roofline_other'].get('kernel_ms_per_step')); print(d['roofline_other'].get('push'))"
PHB_NO_STRIP=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('NO_STRIP', d['value']/1e9, d['ms_per_step']); print(d['roofline_other'].get('kernel_ms_per_step'))"
for k in 1 3 4; do timeout 600 python bench.py --config $k --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('config', $k, d['value']/1e9, d['ms_per_step'], d['roofline_other'].get('kernel_ms_per_step'))"; done
