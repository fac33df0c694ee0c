cd /root/repo
timeout 900 python -m pytest tests/test_host_staging_gpu.py -q 2>&1 | tail -2
timeout 600 python tools/e2e_timeline.py 2>&1 | grep -v Warn | tail -16
timeout 900 python bench.py --no-cpu > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2w.json') if l.startswith('{')][0]
print(d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],2), round(d['python_host']['ms_per_step'],2), round(d['e2e']['value']/1e9,2), d['e2e']['staging_ms_per_rank'])"
