cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --host python --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_r2f_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:tile_kernel' --launch-skip 6 --launch-count 2 -o gpurun_out/prof_r2_final -f python bench.py --host python --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_r2_final.log 2>&1
ls -la gpurun_out/prof_r2_final.ncu-rep gpurun_out/launches_r2f.csv
