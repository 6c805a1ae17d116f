#!/bin/bash
# compute-sanitizer over a representative subset of the GPU tests (memcheck on everything hot; racecheck on the kernels
# that stage through shared memory).  Writes gpurun_out/sanitizer_*.log; summarised under profiles/.
# usage (on the GPU box): bash tools/sanitize.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SUBSET_MEM="tests/test_gpu_set_distance.py tests/test_gpu_knn.py tests/test_gpu_pointnet2.py tests/test_gpu_seams.py tests/test_gpu_edge.py tests/test_gpu_host_step.py"
SUBSET_RACE="tests/test_gpu_knn.py::test_knn_topk_golden tests/test_gpu_knn.py::test_knn_self_vs_oracle tests/test_gpu_knn.py::test_knn_dist_golden tests/test_gpu_knn.py::test_knn_feature_clouds_tensor_core_vs_oracle tests/test_gpu_set_distance.py::test_nn_bidir_golden tests/test_gpu_pointnet2.py tests/test_gpu_edge.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest $SUBSET_MEM -x -q -p no:cacheprovider > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit code: $?" >> gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 0 --error-exitcode 9 \
  python -m pytest $SUBSET_RACE -x -q -p no:cacheprovider > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit code: $?" >> gpurun_out/sanitizer_racecheck.log
# hazards by library: tests/test_gpu_pointnet2.py also runs the REFERENCE's kernels (oracle/_ref/_ext_ref.so) for comparison
python - <<'PY' >> gpurun_out/sanitizer_racecheck.log
import re
txt = open("gpurun_out/sanitizer_racecheck.log", errors="replace").read()
blocks = re.split(r"\n(?==========+ (?:Error|Warning))", txt)
by = {}
for b in blocks:
    if not re.match(r"=+ (Error|Warning)", b):
        continue
    lib = "libhitgeom.so" if "libhitgeom" in b else "_ext_ref.so (reference kernels)" if "_ext_ref" in b else "other"
    m = re.search(r"at (?:void )?([^+(<]+)", b)
    by[(lib, m.group(1).strip() if m else "?")] = by.get((lib, m.group(1).strip() if m else "?"), 0) + 1
print("hazard reports by library / kernel:")
for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
    print(f"  {v:6d}  {k[0]}  {k[1]}")
print("hitgeom hazards:", sum(v for k, v in by.items() if k[0] == "libhitgeom.so"))
PY
tail -5 gpurun_out/sanitizer_memcheck.log; tail -12 gpurun_out/sanitizer_racecheck.log
