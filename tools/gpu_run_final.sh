cd /root/repo
mkdir -p gpurun_out
for k in 1 3; do
timeout 600 python bench.py --config $k --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2q_c$k.json 2> gpurun_out/bench_r2q_c$k.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2q_c$k.json') if l.startswith('{')][0]
print('config $k', d['value']/1e9, d['ms_per_step'], (d['e2e'] or {}).get('value'), d['roofline']['frac'] if d['roofline'] else None, d['roofline_other'].get('kernel_ms_per_step'))"
timeout 600 python bench.py --config $k --steps 10 --warmup 3 --host cpp > gpurun_out/bench_r2q_cpp_c$k.json 2> gpurun_out/bench_r2q_cpp_c$k.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r2q_cpp_c$k.json') if l.startswith('{')][0]
print('cpp config $k', d['value']/1e9, d['ms_per_step'])"
done
