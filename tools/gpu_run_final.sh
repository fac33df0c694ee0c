cd /root/repo
mkdir -p gpurun_out
for tag in py auto py2; do
h=python; [ $tag = auto ] && h=auto
timeout 900 python bench.py --host $h --no-cpu > gpurun_out/bench_r2t_$tag.json 2> gpurun_out/bench_r2t_$tag.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2t_$tag.json') if l.startswith('{')][0]
print('$tag', d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],2), round(d['python_host']['ms_per_step'],2), round(d['e2e']['value']/1e9,2), d['e2e']['staging_ms_per_rank'], d['roofline_other']['kernel_ms_per_step'])"
done
