set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_r2_bench.log 2>&1
tail -2 gpurun_out/launches_r2_bench.log | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'push_tma_kernel|tile_kernel|deposit_scatter_kernel|bin_count_kernel' --launch-skip 9 --launch-count 4 -o gpurun_out/prof_r2_bench python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -8
ls -la gpurun_out/prof_r2_bench.ncu-rep
