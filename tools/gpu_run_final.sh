cd /root/repo
mkdir -p gpurun_out
PHB_MB_ONLY=predict timeout 900 ncu --set full --clock-control none -k regex:tile_kernel --launch-skip 3 --launch-count 3 -o gpurun_out/prof_r2_c4 -f python tools/microbench.py c4 > gpurun_out/prof_r2_c4.log 2>&1
grep "c4 " gpurun_out/prof_r2_c4.log | cut -c1-90
ls -la gpurun_out/prof_r2_c4.ncu-rep
