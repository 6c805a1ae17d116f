#!/bin/bash
# compute-sanitizer over a representative subset of the GPU tests (memcheck on everything hot; racecheck on the kernels
# that stage through shared memory).  Writes gpurun_out/sanitizer_*.log; summarised under profiles/.
# usage (on the GPU box): bash tools/sanitize.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SUBSET_MEM="tests/test_gpu_set_distance.py tests/test_gpu_knn.py tests/test_gpu_pointnet2.py tests/test_gpu_seams.py tests/test_gpu_edge.py tests/test_gpu_host_step.py"
SUBSET_RACE="tests/test_gpu_knn.py::test_knn_topk_golden tests/test_gpu_knn.py::test_knn_self_vs_oracle tests/test_gpu_set_distance.py::test_nn_bidir_golden tests/test_gpu_pointnet2.py tests/test_gpu_edge.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest $SUBSET_MEM -x -q -p no:cacheprovider > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit code: $?" >> gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 \
  python -m pytest $SUBSET_RACE -x -q -p no:cacheprovider > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit code: $?" >> gpurun_out/sanitizer_racecheck.log
tail -5 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log
