cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gridlayout_golden.py tests/test_configs_gpu.py tests/test_amr_gpu.py -q -x 2>&1 | tail -2
timeout 900 python bench.py --no-cpu --no-e2e > gpurun_out/bench_r2ac.json 2> gpurun_out/bench_r2ac.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2ac.json') if l.startswith('{')][0]
print(d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['python_host']['ms_per_step'],3), d['roofline_other']['whole_step_frac_of_hbm'])"
