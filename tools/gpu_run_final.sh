cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predict_gpu.py tests/test_tile_gpu.py -q -x 2>&1 | tail -2
PHB_TILE_BIG=1 timeout 900 python -m pytest tests/test_tile_gpu.py tests/test_configs_gpu.py -q -x 2>&1 | tail -2
timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_solver_gpu.py tests/test_simulator_gpu.py tests/test_amr_gpu.py -q -x 2>&1 | tail -2
for k in 3 1; do
timeout 600 python bench.py --config $k --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2ad_c$k.json 2> gpurun_out/bench_r2ad_c$k.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2ad_c$k.json') if l.startswith('{')][0]
print('config $k', d['host'], round(d['value']/1e9,3), round(d['ms_per_step'],3), round(d['python_host']['ms_per_step'],3), d['roofline_other'].get('kernel_ms_per_step'))"
PHB_PREDICT_MIN_CELLS=0 timeout 600 python bench.py --config $k --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2ad_c${k}_p.json 2> gpurun_out/bench_r2ad_c${k}_p.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2ad_c${k}_p.json') if l.startswith('{')][0]
print('config $k one-pass', d['host'], round(d['value']/1e9,3), round(d['ms_per_step'],3), round(d['python_host']['ms_per_step'],3), d['roofline_other'].get('kernel_ms_per_step'))"
done
