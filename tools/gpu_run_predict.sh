cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo rc=$?
tail -3 gpurun_out/bench_r2n.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2n.json') if l.startswith('{')][0]
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['staging_ms_per_rank']); print(d['roofline_other'].get('kernel_ms_per_step'))"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
