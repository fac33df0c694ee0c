cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_configs_gpu.py tests/test_solver_gpu.py tests/test_cpp_solver.py tests/test_host_api_gpu.py -q -x 2>&1 | tail -2
timeout 900 python bench.py --no-cpu > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2aa.json') if l.startswith('{')][0]
print(d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['python_host']['ms_per_step'],3), round(d['e2e']['value']/1e9,2), d['roofline']['frac'], d['roofline_other']['whole_step_frac_of_hbm'], d['gpu_launches'])"
