cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][0]
print(d['host'], round(d['value']/1e9,2), round(d['ms_per_step'],3), round(d['python_host']['ms_per_step'],3), round(d['e2e']['value']/1e9,2), d['roofline']['frac'], d['roofline_other']['whole_step_frac_of_hbm'], d['cpu_baseline']['value']/1e9, d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
