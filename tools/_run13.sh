set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cd tools
timeout 600 python prof_strip.py c5 2>&1 | tail -5
timeout 600 python prof_strip.py c3 2>&1 | tail -5
timeout 600 python prof_strip.py c1 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"push_strip_kernel|push_tma_kernel" --launch-skip 4 --launch-count 2 -o ../gpurun_out/prof_strip_r2 python prof_strip.py c5s 2>&1 | tail -4
