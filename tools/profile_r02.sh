#!/bin/bash
# Round-2 profile captures (run on the GPU box: bash tools/profile_r02.sh).  Writes gpurun_out/r2_*; the summaries kept
# under profiles/ are produced from these by tools/ncu_summary.py / tools/launch_summary.py in the build container.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch lists (every launch with its device time; cold-cache, serialised: compare SHARES)
$NCU --metrics gpu__time_duration.sum -s 0 -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-subrecords > gpurun_out/r2_launches_c5.log 2>&1
$NCU --metrics gpu__time_duration.sum -s 0 -c 1500 --csv --log-file gpurun_out/r2_launches_c1.csv \
  python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_c1.log 2>&1
# one --set full capture per hot kernel
$NCU --set full --import-source on -k regex:knn3_small_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_knn_small \
  python tools/prof_one.py knn 388 1024 > gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:knn_seed_small -s 2 -c 1 -f -o gpurun_out/r2_prof_knn_seed_small \
  python tools/prof_one.py knn 388 1024 >> gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:nn_bidir_d3_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_nn_1024x16384 \
  python tools/prof_one.py nn 1024 16384 0 0 2 >> gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:knn3_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_knn_1024x16384 \
  python tools/prof_one.py knn 1024 16384 0 0 2 >> gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:nn_bidir_d3_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_nn_388x1024 \
  python tools/prof_one.py nn 388 1024 >> gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:knn_tc_fused -s 2 -c 1 -f -o gpurun_out/r2_prof_knn_tc_fused \
  python tools/debug/prof_knn_tc.py 64 >> gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:scatter_bulk -s 2 -c 1 -f -o gpurun_out/r2_prof_scatter_bulk \
  python tools/debug/prof_scatter.py 64 >> gpurun_out/r2_ncu.log 2>&1
$NCU --set full --import-source on -k regex:csr_build_smem -s 2 -c 1 -f -o gpurun_out/r2_prof_csr_build \
  python tools/debug/prof_scatter.py 64 >> gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
ls -la gpurun_out/r2_prof_*.ncu-rep | awk '{print $5, $9}'
